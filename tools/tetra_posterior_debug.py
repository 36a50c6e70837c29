import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np
from instruct_b200 import Sampler, SeqData, _lib
g = np.load('/root/repo/tests/golden/posterior_tetra.npz')
K = int(g["K"]); R = g["S"].shape[0]
update, burnin, thinning = int(g["update"]), int(g["burnin"]), int(g["thinning"])
sd = SeqData(g["x"], g["allelenum"], K, ploid=4, autopoly=1)
pop = g["pop"]
print("update", update, burnin, thinning, "refS", np.round(g["S"], 3).tolist(), "refLL", np.round(g["LL"]).tolist())
def run(rep, seed):
    s = Sampler(sd, seed=seed)
    s.chain_init(rep, initd=[0.3 + 0.02 * rep, 0.6 - 0.02 * rep])
    s.set(_lib.STATE_ALPHA, [float(g["alpha"][rep])])
    accS, accQ, accL, n = np.zeros(K), np.zeros((s.N, K)), 0.0, 0
    tr = []
    for step in range(update):
        s.sweep(1)
        if step >= burnin and (step + 1 - burnin) % thinning == 0:
            accS += s.get(_lib.STATE_S); accQ += s.get(_lib.STATE_Q); l = float(s.get(_lib.STATE_TOTALLKH)[0]); accL += l; n += 1; tr.append(l)
    s.close()
    S, Q, LL = accS / n, accQ / n, accL / n
    o = np.argsort(Q[pop == 0].mean(axis=0))[::-1]
    return S[o], LL, Q[pop == 0][:, o[0]].mean(), tr
for rep in range(R):
    S, LL, q, tr = run(rep, 2000 + rep)
    print(rep, "S", np.round(S, 3), "dS", np.round(S - g["S"][rep], 3), "LL", round(LL), "dLL", round(LL - g["LL"][rep]), "q", round(q, 3), "trace", [round(v) for v in tr[::max(1, len(tr)//6)]])
    if abs(LL - g["LL"][rep]) > 500:
        for sd2 in (3000 + rep, 4000 + rep, 5000 + rep):
            S, LL, q, tr = run(rep, sd2)
            print("   retry seed", sd2, "S", np.round(S, 3), "dLL", round(LL - g["LL"][rep]))
