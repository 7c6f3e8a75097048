"""Numerical study for DESIGN.md section 8.5: FP64 GEMM emulated with radix-256 integer digit products (the arithmetic an int8
tcgen05 path would run): error against an 80-bit reference on a C4-like ill-conditioned K(X,Z) * (V~ P), for several digit counts.
CPU only (numpy); not part of the product or the tests."""
import numpy as np, math, time
rng = np.random.default_rng(0)
n, d, m, j = 2048, 8, 1024, 32
x = rng.standard_normal((n + m, d))
ls = np.array([math.sqrt(d) * (0.75 + 0.5 * k / (d - 1)) for k in range(d)])
def rbf(a, b):
    a = a / ls; b = b / ls
    d2 = (a * a).sum(1)[:, None] + (b * b).sum(1)[None] - 2 * a @ b.T
    return np.exp(-0.5 * np.maximum(d2, 0))
z = x[:m]; xs = x[m:]
kzz = rbf(z, z)
lam, vec = np.linalg.eigh(kzz / m)
keep = lam > 0
vt = vec[:, keep] / np.sqrt(keep.sum() * lam[keep])
p = rng.standard_normal((keep.sum(), j))
w = vt @ p
k = rbf(xs, z)
print("W max", np.abs(w).max(), "lam min", lam[keep].min())
truth = (k.astype(np.longdouble) @ w.astype(np.longdouble))
native = k @ w
scale = np.abs(truth).max()
print("F max", float(scale), "native err", float(np.abs(native - truth).max() / scale))

def digits_unsigned(a, nd):  # a in [0, 1): radix-256 floor digits; returns list of uint8 arrays, a ~= sum d_i 256^-(i+1)
    out = []; r = a.copy()
    for _ in range(nd):
        r = r * 256.0
        dgt = np.floor(r); r = r - dgt
        out.append(dgt.astype(np.int64))
    return out
def digits_signed(b, nd):  # b in [-1, 1): top digit signed (floor(b*128)), rest unsigned; b ~= d0/128 + sum_{i>=1} d_i /(128*256^i)
    out = []; r = b * 128.0
    dgt = np.floor(r); r = r - dgt; out.append(dgt.astype(np.int64))
    for _ in range(nd - 1):
        r = r * 256.0
        dgt = np.floor(r); r = r - dgt
        out.append(dgt.astype(np.int64))
    return out

for na, nb, tmax in [(7, 7, 6), (7, 8, 7), (8, 8, 7), (6, 6, 5)]:
    ka = digits_unsigned(k * 0.5, na)
    e = np.ceil(np.log2(np.abs(w).max(0))) + 1  # |w / 2^e| <= 0.5
    wb = digits_signed(w / 2.0 ** e, nb)
    acc = np.zeros((n, j), dtype=np.longdouble)
    cnt = 0
    for t in range(tmax + 1):
        g = np.zeros((n, j), dtype=np.int64)
        for i in range(min(t, na - 1) + 1):
            jj = t - i
            if jj < nb:
                g += ka[i] @ wb[jj]; cnt += 1
        assert np.abs(g).max() < 2**31, np.abs(g).max()
        acc += g.astype(np.longdouble) * np.longdouble(2.0) ** (-8 * (t + 1) - 7)
    res = (acc * 2 * (2.0 ** e)[None]).astype(np.float64)
    print(na, nb, tmax, "pairs", cnt, "err", float(np.abs(res - truth).max() / scale))
