"""Planning aid for the sharded path: the ladder gather stays of the mask form in both layouts,
layout A moves no data between ranks, layout B has exactly one cross-rank CNOT -- checked numerically on the bit matrices."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))


def test_two_layout_ladder_structure():
    import shard_layout_model as m
    for n in range(8, 14):
        for g in (1, 2, 3):
            if n - g < 2 * g + 2:
                continue
            for stacking in (0, 1):
                a = m.analyse(n, g, stacking, "A")
                b = m.analyse(n, g, stacking, "B")
                assert a["Mc"] == 0                                   # natural layout: relabelling + local gather only
                assert b["Mc"] != 0 and b["Mc"] >> (n - 2 * g)        # swapped layout: control(s) inside the top g local bits
                assert bin(b["Mc"]).count("1") <= g
