"""Tier-2 parity harness: do two ensembles of integer trajectories come from the same law?

Per (sample time, species): two-sample Kolmogorov-Smirnov at family-wise level ALPHA (Bonferroni
over all rows), |mean_a - mean_b| <= Z * standard error, and a variance-ratio band from the
fourth-moment standard error.  n, m, ALPHA and Z are stated by the caller and printed on failure
(the tolerance is part of the test).
"""
import numpy as np


def ks_statistic(a, b):
    """sup |F_a - F_b| for two integer samples (exact, ties handled by evaluating after each distinct value)."""
    a = np.sort(np.asarray(a))
    b = np.sort(np.asarray(b))
    grid = np.union1d(a, b)
    fa = np.searchsorted(a, grid, side="right") / a.size
    fb = np.searchsorted(b, grid, side="right") / b.size
    return float(np.max(np.abs(fa - fb)))


def ks_critical(n, m, alpha):
    """Asymptotic two-sample critical value: D = sqrt(-ln(alpha/2)/2) * sqrt((n+m)/(n*m))."""
    return float(np.sqrt(-np.log(alpha / 2.0) / 2.0) * np.sqrt((n + m) / (n * m)))


def compare_ensembles(a, b, alpha=1e-3, z=5.0):
    """a: [rows..., n], b: [rows..., m] integer samples.  Returns a list of failure strings (empty = same law)."""
    a = np.asarray(a).reshape(-1, np.asarray(a).shape[-1]).astype(np.float64)
    b = np.asarray(b).reshape(-1, np.asarray(b).shape[-1]).astype(np.float64)
    rows, n = a.shape
    m = b.shape[1]
    crit = ks_critical(n, m, alpha / rows)
    fails = []
    for r in range(rows):
        x, y = a[r], b[r]
        d = ks_statistic(x, y)
        if d > crit:
            fails.append(f"row {r}: KS D={d:.4f} > {crit:.4f} (n={n}, m={m}, alpha={alpha}/{rows})")
        vx, vy = x.var(ddof=1) if n > 1 else 0.0, y.var(ddof=1) if m > 1 else 0.0
        se = np.sqrt(vx / n + vy / m)
        if abs(x.mean() - y.mean()) > z * se + 1e-12:
            fails.append(f"row {r}: means {x.mean():.4f} vs {y.mean():.4f} differ by more than {z} se ({se:.4g})")
        if vx > 0 and vy > 0:
            # standard error of a sample variance: sqrt((mu4 - var^2) / n)
            mu4x, mu4y = np.mean((x - x.mean()) ** 4), np.mean((y - y.mean()) ** 4)
            sev = np.sqrt(max(mu4x - vx * vx, 0) / n + max(mu4y - vy * vy, 0) / m)
            if abs(vx - vy) > z * sev + 1e-12:
                fails.append(f"row {r}: variances {vx:.4f} vs {vy:.4f} differ by more than {z} se ({sev:.4g})")
    return fails
