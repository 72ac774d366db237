"""How often would a RE-ASSOCIATED donut sum change a decision?  (CPU study for the next round)

The score kernel is bound by the reference's one-add-at-a-time fp64 order.  A faster kernel could sum in another order
(ring / row partial sums shared between pixels) if every DECISION taken on the result is provably unchanged: lambda-chunk
membership (strict edges), E > 0, E.max() -> number of chunks; the few candidates / survivors would be re-evaluated in
the exact order.  This script measures, on the bench generator, the distance between exact and re-associated sums and
how close expected values come to a chunk edge."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from hicpeaks_b200.synth import synth_chromosome
from oracle import hiccups_oracle as ho

n, band, pw, ww, maxww, thr = int(sys.argv[1]) if len(sys.argv) > 1 else 2500, 500, [2], [5], 10, 16
inp = synth_chromosome(n, band, 5, maxww=maxww, seed=17)
t = time.time()
sw = ho.sweep(inp, pw, ww, maxww, thr, band)
print("exact sweep %.1f s, pixels %d, frozen %d" % (time.time() - t, sw["total"], sw["frozen"]))
raw, bal, eb = ho.dense_band(inp)
num = inp["num"]
vx, vd = sw["vx"], sw["vd"]
res_w = sw["res_w"][2]
prog = ho.step_program(pw, ww, maxww)
alt = [np.zeros((num, n)), np.zeros((num, n))]
altV = [np.zeros(sw["total"]), np.zeros(sw["total"])]
for s in sw["executed"]:
    p, w, ops = prog[s]
    for fl in (0, 1):
        parts = []
        for a, b, is_y, _ in ops:
            if fl == 1 and not is_y:
                continue
            z = np.zeros((num, n))
            ho._add_shift(z, bal, a, b)
            parts.append(z)
        # pairwise tree over the cells of the step, then one add into the running total: a different association
        while len(parts) > 1:
            parts = [parts[i] + parts[i + 1] if i + 1 < len(parts) else parts[i] for i in range(0, len(parts), 2)]
        alt[fl] += parts[0]
        hit = np.nonzero(res_w == w)[0]
        altV[fl][hit] = alt[fl][vd[hit], vx[hit]]
B = inp["biases"]
ir = np.zeros(num)
for d, v in inp["IR"].items():
    ir[d] = v
edges = np.array([0.0] + [e[1] for e in ho.chunk_edges(60)])
for fl, name in ((0, "donut"), (1, "lower-left")):
    bs, be = sw["bSV"][2][fl], sw["bEV"][2][fl]
    m = (be != 0) & (res_w > 0)
    x, d = vx[m], vd[m]
    E = ir[d] * (bs[m] / be[m]) * B[x] * B[x + d]
    Ea = ir[d] * (altV[fl][m] / be[m]) * B[x] * B[x + d]
    pos = E > 0
    rel = np.abs(Ea[pos] - E[pos]) / E[pos]
    ch = np.searchsorted(edges, E[pos], side="left")
    cha = np.searchsorted(edges, Ea[pos], side="left")
    k = np.searchsorted(edges, E[pos])
    lo, hi = edges[np.maximum(k - 1, 0)], edges[np.minimum(k, edges.size - 1)]
    gap = np.minimum(np.abs(E[pos] - lo) / E[pos], np.abs(hi - E[pos]) / E[pos])
    print("%-10s pixels %d | sums differ in %d (%.1f %%), max rel diff of E %.2e | zero <-> non-zero flips %d | chunk changes %d | "
          "closest approach to an edge (relative) %.2e, within 1e-12: %d, within 1e-9: %d" % (
              name, pos.sum(), int((Ea[pos] != E[pos]).sum()), 100.0 * (Ea[pos] != E[pos]).mean(), rel.max(),
              int(((Ea > 0) != (E > 0)).sum()), int((ch != cha).sum()), gap.min(), int((gap < 1e-12).sum()), int((gap < 1e-9).sum())))
