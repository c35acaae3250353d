"""Hyper-parameter proposals for Synthesize.adaptive (reference main.py:763-810 uses hyperopt's TPE
over `num_cp_gates ~ quniform(min, max, 1)` and `r ~ lognormal(log r_mean, r_variance)`).

hyperopt is not a dependency here; `TPESampler` is a small tree-structured Parzen estimator over
the same two-dimensional space: random draws from the prior for the first `n_startup` evaluations,
afterwards candidates are drawn from a Parzen mixture fitted to the best `gamma` fraction of the
trials and the one maximising l(x)/g(x) is proposed.  `next_seed` reproduces the reference's seed
chain (`_, subkey = random.split(PRNGKey(seed)); seed = int(subkey[1])`, main.py:798-799) with
jax 0.3.x threefry semantics — integer work on the host, no device involved.
"""
import math

import numpy as np

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))
_M32 = 0xFFFFFFFF


def _threefry2x32(k0, k1, x0, x1):
    ks = (k0, k1, k0 ^ k1 ^ 0x1BD11BDA)
    x0 = (x0 + ks[0]) & _M32
    x1 = (x1 + ks[1]) & _M32
    for g in range(5):
        for r in _ROT[g % 2]:
            x0 = (x0 + x1) & _M32
            x1 = ((x1 << r) | (x1 >> (32 - r))) & _M32
            x1 ^= x0
        x0 = (x0 + ks[(g + 1) % 3]) & _M32
        x1 = (x1 + ks[(g + 2) % 3] + g + 1) & _M32
    return x0, x1


def next_seed(seed):
    """int(split(PRNGKey(seed))[1][1]): split hashes iota(4) as the pairs (0,2), (1,3); the wanted
    word is the second output of the pair (1, 3)."""
    seed = int(seed)
    k0, k1 = (seed >> 32) & _M32, seed & _M32
    return _threefry2x32(k0, k1, 1, 3)[1]


class TPESampler:
    def __init__(self, min_num_cp_gates, max_num_cp_gates, r_mean, r_variance, n_startup=20, gamma=0.25,
                 n_candidates=24):
        self.lo, self.hi = int(min_num_cp_gates), int(max_num_cp_gates)
        self.mu, self.sigma = math.log(r_mean), float(r_variance)
        self.n_startup, self.gamma, self.n_candidates = n_startup, gamma, n_candidates

    def _prior(self, rng):
        k = int(round(rng.uniform(self.lo - 0.5, self.hi + 0.5)))
        k = min(max(k, self.lo), self.hi)
        return k, float(math.exp(rng.normal(self.mu, self.sigma)))

    def _log_parzen(self, x, pts, bw, prior_mu, prior_sigma):
        """log density at x [m] of an equal-weight mixture of N(pt, bw) plus the prior component."""
        comps = [-0.5 * ((x - prior_mu) / prior_sigma) ** 2 - math.log(prior_sigma)]
        for p in pts:
            comps.append(-0.5 * ((x - p) / bw) ** 2 - math.log(bw))
        c = np.stack(comps)
        mx = c.max(0)
        return mx + np.log(np.exp(c - mx).sum(0)) - math.log(len(comps))

    def suggest(self, results, rng):
        """Next (num_cp_gates, r) given the list of finished result dicts (keys num_cp_gates, r, loss)."""
        done = [t for t in results if math.isfinite(t['loss'])]
        if len(results) < self.n_startup or len(done) < 4:
            return self._prior(rng)
        done = sorted(done, key=lambda t: t['loss'])
        n_good = max(2, int(math.ceil(self.gamma * len(done))))
        good, bad = done[:n_good], done[n_good:] + [t for t in results if not math.isfinite(t['loss'])]
        span = max(1.0, self.hi - self.lo)
        kb, rb = max(1.0, span / 6), max(0.1, self.sigma / 2)
        gk = np.array([t['num_cp_gates'] for t in good], float)
        gr = np.log([t['r'] for t in good])
        bk = np.array([t['num_cp_gates'] for t in bad], float)
        br = np.log([t['r'] for t in bad]) if bad else np.zeros(0)
        # candidates from the "good" mixture (prior included as one component)
        cand_k, cand_r = [], []
        for _ in range(self.n_candidates):
            j = rng.integers(0, len(good) + 1)
            if j == len(good):
                k, r = self._prior(rng)
                cand_k.append(k)
                cand_r.append(math.log(r))
            else:
                cand_k.append(min(max(int(round(rng.normal(gk[j], kb))), self.lo), self.hi))
                cand_r.append(rng.normal(gr[j], rb))
        ck, cr = np.array(cand_k, float), np.array(cand_r, float)
        mid, wid = (self.lo + self.hi) / 2, span
        score = (self._log_parzen(ck, gk, kb, mid, wid) + self._log_parzen(cr, gr, rb, self.mu, self.sigma)
                 - self._log_parzen(ck, bk, kb, mid, wid) - self._log_parzen(cr, br, rb, self.mu, self.sigma))
        best = int(np.argmax(score))
        return int(ck[best]), float(math.exp(cr[best]))
