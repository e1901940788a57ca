"""Plain-Python restatement of SpeciesMetrics::calculateAbundance's iteration (aln_sink.h:196-495) on a flattened
tie-set table, and a seeded generator of such tables.  Test infrastructure: Python floats are IEEE doubles and the
loops add in the reference's order, so results are comparable bit for bit."""
import math

import numpy as np


def em_python(count, key_off, target, length, p):
    """SpeciesMetrics::calculateAbundance iteration (aln_sink.h:410-480) on the flattened table."""
    n, K = len(p), len(count)

    def step(p):
        pn = [0.0] * n
        for k in range(K):
            tg = target[key_off[k]:key_off[k + 1]]
            psum = 0.0
            for j in tg:
                psum += p[j]
            if psum == 0.0:
                continue
            for j in tg:
                pn[j] += count[k] * (p[j] / psum)
        s = 0.0
        for i in range(n):
            s += pn[i] / length[i]
        return [pn[i] / length[i] / s for i in range(n)]

    it = 0
    while True:
        pn = step(p)
        pn2 = step(pn)
        ssr = ssv = 0.0
        pr = [0.0] * n; pv = [0.0] * n
        for i in range(n):
            pr[i] = pn[i] - p[i]; ssr += pr[i] * pr[i]
            pv[i] = pn2[i] - pn[i] - pr[i]; ssv += pv[i] * pv[i]
        if ssv > 0.0:
            g = -math.sqrt(ssr / ssv)
            pn2 = [max(0.0, p[i] - 2 * g * pr[i] + g * g * pv[i]) for i in range(n)]
            pn = step(pn2)
        diff = 0.0
        for i in range(n):
            diff += (p[i] - pn[i]) if p[i] > pn[i] else (pn[i] - p[i])
        if diff < 0.0000000001:
            break
        it += 1
        if it >= 10000:
            break
        p = pn
    return p, it, diff



def random_problem(seed, n, K):
    rng = np.random.default_rng(seed)
    count, key_off, target = [], [0], []
    for _ in range(K):
        sz = int(rng.integers(1, 5))
        ids = rng.integers(0, n, size=sz)              # duplicates allowed: an ancestor and its own leaf in one key
        target += [int(x) for x in ids]
        count.append(int(rng.integers(1, 5000)))
        key_off.append(len(target))
    length = [int(x) for x in rng.integers(1000, 5_000_000, size=n)]
    if n > 6:
        length[3] = 2 ** 64 - 1                          # "no size known" (numeric_limits<size_t>::max())
    p0 = rng.random(n); p0[rng.random(n) < 0.15] = 0.0   # species nobody hit
    if p0.sum() == 0:
        p0[0] = 1.0
    p0 = [float(x) for x in p0 / p0.sum()]
    return count, key_off, target, length, p0
