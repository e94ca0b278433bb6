"""Build checks that need no GPU: the library cross-compiles for sm_100a and the hot tile kernels
(McClean instantiations of k_tile12) fit their register budget without local-memory spills --
a spill in the persistent tile loop costs ~5 % of a 30-qubit backward launch (profiles/README.md)."""
import re
import shutil
import subprocess

import pytest

from qradient_b200 import build as qbuild


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not available")
def test_hot_tile_kernels_compile_without_spills(tmp_path):
    import os
    out = tmp_path / "lib_check.so"
    cmd = [qbuild.nvcc_path()] + qbuild.NVCC_FLAGS + ["-Xptxas", "-v"] + [os.path.join(qbuild.CSRC, f) for f in qbuild.SOURCES] + ["-o", str(out)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
    log = res.stdout + res.stderr
    # ptxas prints "Function properties for <mangled>" followed by the stack / spill line and the register line
    blocks = re.findall(r"Function properties for (\S+)\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n"
                        r"ptxas info\s+: Used (\d+) registers", log)
    # PHASE = false (McClean passes)
    hot = [b for b in blocks if b[0].startswith("_Z8k_tile12") and re.search(r"k_tile12ILi\dELb0E", b[0])]
    assert len(hot) >= 6, [b[0] for b in blocks][:20]
    for name, stack, st, ld, regs in hot:
        assert int(st) == 0 and int(ld) == 0, (name, st, ld)
        nv = int(re.search(r"k_tile12ILi(\d)E", name).group(1))
        k11 = "ELi11EEv8TilePass" in name
        staged = int(re.search(r"ELb0ELi(\d)E", name).group(1))
        limit = 64 if (nv == 1 and not staged) else 128     # forward: 2 x 512 (or 4 x 256) threads per SM; backward: 512 (2 x 256)
        assert int(regs) <= limit, (name, regs, k11)
    # the exchange-pass instantiations of sharded registers (k_tile12_x<NV, PHASE = false>): same budgets, no spills
    xhot = [b for b in blocks if b[0].startswith("_Z10k_tile12_xILi") and "ELb0E" in b[0]]
    assert len(xhot) == 2, [b[0] for b in blocks if "tile12_x" in b[0]]
    for name, stack, st, ld, regs in xhot:
        assert int(st) == 0 and int(ld) == 0, (name, st, ld)
        assert int(regs) <= (64 if "ILi1E" in name else 128), (name, regs)
