"""Golden table for the t-test p-value: 2 * Q(t, nu) = I_{nu/(nu+t^2)}(nu/2, 1/2), the regularised
incomplete beta function, evaluated with mpmath at 50 digits.  This is what the reference obtains as
2 * gsl_cdf_tdist_Q(t, nu) (src/setComparisons.c:117); GSL is not in this image, so the table -- not
the oracle's own implementation -- is what the DEVICE function is pinned on.
  python tests/golden/make_tdist_golden.py  ->  tests/golden/tdist_table.json"""
import json
import os

import mpmath as mp

mp.mp.dps = 50
rows = []
ts = [0.0, 1e-8, 1e-3, 0.05, 0.3, 0.7, 1.0, 1.5, 2.0, 2.5, 3.3, 5.0, 8.0, 13.0, 25.0, 60.0, 200.0, 1e3, 1e5]
nus = [1.0, 1.5, 2.0, 2.7, 3.0, 4.0, 5.5, 8.0, 13.0, 20.0, 37.5, 64.0, 100.0, 333.3, 1e3, 1e4, 1e5]
for nu in nus:
    for t in ts:
        x = mp.mpf(nu) / (mp.mpf(nu) + mp.mpf(t) ** 2)
        p2 = mp.betainc(mp.mpf(nu) / 2, mp.mpf(1) / 2, 0, x, regularized=True)      # two-sided p = 2 Q(t, nu)
        rows.append([t, nu, float(p2)])
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "tdist_table.json"), "w") as f:
    json.dump({"doc": "two-sided Student-t p-value 2*Q(t,nu), mpmath 50 digits; columns t, nu, p", "rows": rows}, f)
print(len(rows), "rows")
