"""Index algebra of the two-layout schedule for sharded states -- a planning aid, not product code.

Layout A: physical index (rank r, local l) holds the logical amplitude j = r << n_loc | l  (qubits 0..g-1 on the rank bits).
Layout B: rank bits and the top g local bits have traded places (what k_global_gates produces when it stores its results
locally): qubits g..2g-1 on the rank bits, qubits 0..g-1 on the top g local bits.

The CNOT-ladder gather is GF(2)-linear, new[j] = old[L(j)], L(j) = j ^ ((j >> 1) & M1) ^ ((j >> 2) & M2).  In a layout with
logical -> physical index map S the gather reads  F = S o L o S^-1.  This script computes F as a bit matrix for both
layouts and both stackings and checks the structural claims the kernels rely on:

  (1) local part:  src_local = l ^ ((l >> 1) & M1') ^ ((l >> 2) & M2') ^ X(rank)      (mask form + a rank-dependent constant)
  (2) rank part:   src_rank  = R(rank) ^ (parity(l & Mc) ? E : 0)                      (Mc inside the top g local bits)

and prints M1', M2', Mc, E.  Layout A must give Mc = 0 (no data crosses ranks: the ladder is a relabelling plus a local
gather); layout B gives exactly one local-control -> global-target CNOT, CNOT(g-1 -> g).

    python scripts/shard_layout_model.py [n] [g]
"""
import sys


def ladder_masks(n, stacking):
    """scatter masks of ladder `stacking` (state.py:229-241; same rule as qr_lib.cu ladder_masks)"""
    a = b = 0
    for t in range(1, n):
        p = n - 1 - t
        a |= 1 << p
        three = (t % 2 == 1) if stacking == 0 else (t % 2 == 0)
        if t >= 2 and three:
            b |= 1 << p
    return a, b


def lmap(j, m1, m2):
    return j ^ ((j >> 1) & m1) ^ ((j >> 2) & m2)


def swap_blocks(j, n, g):
    """layout B: exchange the bit blocks [n-1 .. n-g] and [n-g-1 .. n-2g] (an involution)"""
    n_loc = n - g
    top = (j >> n_loc) & ((1 << g) - 1)
    mid = (j >> (n_loc - g)) & ((1 << g) - 1)
    low = j & ((1 << (n_loc - g)) - 1)
    return (mid << n_loc) | (top << (n_loc - g)) | low


def analyse(n, g, stacking, layout):
    n_loc = n - g
    # the GATHER of ladder(stacking) uses the masks of the inverse ladder (qr_lib.cu launch_ladder)
    m1, m2 = ladder_masks(n, 1 - stacking)
    S = (lambda j: j) if layout == "A" else (lambda j: swap_blocks(j, n, g))
    F = lambda p: S(lmap(S(p), m1, m2))
    cols = [F(1 << k) for k in range(n)]
    # linearity
    for p in (0x155 & ((1 << n) - 1), (1 << n) - 1, 0x2b3 & ((1 << n) - 1)):
        acc = 0
        for k in range(n):
            if (p >> k) & 1:
                acc ^= cols[k]
        assert acc == F(p), "map is not linear"
    lmask = (1 << n_loc) - 1
    M1 = M2 = 0
    Mc, E = 0, None
    for k in range(n_loc):                      # local input bits
        c = cols[k]
        loc, rk = c & lmask, c >> n_loc
        assert (loc >> k) & 1, "diagonal missing"
        rest = loc ^ (1 << k)
        if k >= 1 and (rest >> (k - 1)) & 1:
            M1 |= 1 << (k - 1); rest ^= 1 << (k - 1)
        if k >= 2 and (rest >> (k - 2)) & 1:
            M2 |= 1 << (k - 2); rest ^= 1 << (k - 2)
        assert rest == 0, "local part is not of the mask form (layout %s, bit %d: %s)" % (layout, k, bin(c))
        if rk:
            Mc |= 1 << k
            assert E is None or E == rk, "more than one rank pattern"
            E = rk
    X = [cols[n_loc + b] & lmask for b in range(g)]            # rank bit b -> XOR constant on the local source index
    R = [cols[n_loc + b] >> n_loc for b in range(g)]           # rank bit b -> source rank bits
    if Mc:
        assert Mc >> (n_loc - g), "control of the cross-rank CNOT is not in the top g local bits"
    return dict(M1=M1, M2=M2, Mc=Mc, E=E or 0, X=X, R=R)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    g = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    cases = [(n, g)] if n else [(nn, gg) for nn in range(8, 15) for gg in (1, 2, 3) if nn - gg >= 2 * gg + 2]
    for nn, gg in cases:
        for stacking in (0, 1):
            a = analyse(nn, gg, stacking, "A")
            b = analyse(nn, gg, stacking, "B")
            assert a["Mc"] == 0, "layout A must not move data between ranks"
            n_loc = nn - gg
            # layout B: the local masks are those of layout A with the adjacency (qubit g-1 | qubit 2g) removed, restricted
            # to the local bits; exactly one local bit (that of qubit g-1, the lowest of the top block) may feed the rank bits
            print("n=%2d g=%d stacking=%d | A: M1=%s M2=%s X=%s R=%s | B: M1=%s M2=%s Mc=%s E=%s X=%s R=%s" % (
                nn, gg, stacking, bin(a["M1"]), bin(a["M2"]), [bin(x) for x in a["X"]], [bin(r) for r in a["R"]],
                bin(b["M1"]), bin(b["M2"]), bin(b["Mc"]), bin(b["E"]), [bin(x) for x in b["X"]], [bin(r) for r in b["R"]]))
            assert bin(b["Mc"]).count("1") <= gg
    print("all structural checks passed")


if __name__ == "__main__":
    main()
