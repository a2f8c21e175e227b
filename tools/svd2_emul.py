"""Emulation of LAPACK dgesdd on a 2x2 real matrix (dgebrd Householder + dbdsqr/dlasv2), to pin the
sign conventions numpy.linalg.svd returns."""
import numpy as np
EPS = np.finfo(float).eps / 2          # dlamch('E') = 1.1e-16
UNFL = np.finfo(float).tiny
def sign(a, b):
    return abs(a) if (b > 0 or (b == 0 and not np.signbit(b))) else -abs(a)

def dlasv2(f, g, h):
    ft, fa, ht, ha = f, abs(f), h, abs(h)
    pmax = 1
    swap = ha > fa
    if swap:
        pmax = 3
        ft, ht = ht, ft
        fa, ha = ha, fa
    gt, ga = g, abs(g)
    if ga == 0:
        ssmin, ssmax, clt, crt, slt, srt = ha, fa, 1., 1., 0., 0.
    else:
        gasmal = True
        if ga > fa:
            pmax = 2
            if fa / ga < EPS:
                gasmal = False
                ssmax = ga
                ssmin = fa / (ga / ha) if ha > 1 else (fa / ga) * ha
                clt = 1.; slt = ht / gt; srt = 1.; crt = ft / gt
        if gasmal:
            d = fa - ha
            l = 1. if d == fa else d / fa
            m = gt / ft
            t = 2. - l
            mm, tt = m * m, t * t
            s = np.sqrt(tt + mm)
            r = abs(m) if l == 0 else np.sqrt(l * l + mm)
            a = 0.5 * (s + r)
            ssmin, ssmax = ha / a, fa * a
            if mm == 0:
                if l == 0:
                    t = sign(2., ft) * sign(1., gt)
                else:
                    t = gt / sign(d, ft) + m / t
            else:
                t = (m / (s + t) + m / (r + l)) * (1. + a)
            l = np.sqrt(t * t + 4.)
            crt, srt = 2. / l, t / l
            clt = (crt + srt * m) / a
            slt = (ht / ft) * srt / a
    if swap:
        csl, snl, csr, snr = srt, crt, slt, clt
    else:
        csl, snl, csr, snr = clt, slt, crt, srt
    if pmax == 1:
        tsign = sign(1., csr) * sign(1., csl) * sign(1., f)
    elif pmax == 2:
        tsign = sign(1., snr) * sign(1., csl) * sign(1., g)
    else:
        tsign = sign(1., snr) * sign(1., snl) * sign(1., h)
    ssmax = sign(ssmax, tsign)
    ssmin = sign(ssmin, tsign * sign(1., f) * sign(1., h))
    return ssmin, ssmax, snr, csr, snl, csl

def svd2(A):
    a, b, c, d = A[0, 0], A[0, 1], A[1, 0], A[1, 1]
    # dgebrd: H1 zeroes A[1,0]
    if c == 0:
        Q = np.eye(2); d1, e, d2 = a, b, d
    else:
        beta = -sign(np.hypot(a, c), a)
        tau = (beta - a) / beta
        v = c / (a - beta)
        w = b + v * d                       # [1 v] . column 2
        e = b - tau * w
        d2 = d - tau * v * w
        d1 = beta
        Q = np.eye(2) - tau * np.array([[1, v], [v, v * v]])
    # dbdsdc scales by the max-norm, dbdsqr deflates negligible e
    nrm = max(abs(d1), abs(d2), abs(e))
    U = np.eye(2); VT = np.eye(2)
    if nrm == 0:
        s = np.array([0., 0.])
    else:
        d1s, d2s, es = d1 / nrm, d2 / nrm, e / nrm
        tolmul = max(10., min(100., EPS ** (-0.125)))
        tol = tolmul * EPS
        smin = abs(d1s)
        if smin != 0:
            mu = abs(d2s) * (smin / (smin + abs(es)))
            smin = min(smin, mu)
        smin /= np.sqrt(2.)
        thresh = max(tol * smin, 6 * 2 * 2 * UNFL)
        if abs(es) <= thresh:
            sv = [d1s, d2s]
        else:
            ssmin, ssmax, snr, csr, snl, csl = dlasv2(d1s, es, d2s)
            sv = [ssmax, ssmin]
            VT = np.array([[csr, snr], [-snr, csr]])
            U = np.array([[csl, -snl], [snl, csl]])
        for i in range(2):
            if sv[i] < 0 or (sv[i] == 0 and np.signbit(sv[i])):
                sv[i] = -sv[i]
                VT[i] = -VT[i]
        if sv[0] < sv[1]:
            sv = sv[::-1]; U = U[:, ::-1]; VT = VT[::-1]
        s = np.array(sv) * nrm
    return Q @ U, s, VT

if __name__ == "__main__":
    rng = np.random.default_rng(1)
    cases = [np.eye(2) + 0.3 * rng.normal(size=(2, 2)) for _ in range(20000)]
    cases += [rng.normal(size=(2, 2)) * 10 ** rng.uniform(-3, 3) for _ in range(20000)]
    for _ in range(2000):                       # upper triangular / diagonal / rotations / singular
        a, b, d = rng.normal(size=3)
        cases += [np.array([[a, b], [0, d]]), np.array([[a, 0], [0, d]]), np.array([[a, 1e-17 * b], [0, d]])]
        t = rng.uniform(-np.pi, np.pi)
        R = np.array([[np.cos(t), -np.sin(t)], [np.sin(t), np.cos(t)]])
        cases += [R, a * R, R @ np.diag([a, d]), np.outer(rng.normal(size=2), rng.normal(size=2))]
    cases += [np.eye(2), np.zeros((2, 2)), -np.eye(2), np.array([[0, 1.], [1, 0]]), np.array([[1., 0], [0, -1]]),
              np.array([[0, 0], [0, 1.]]), np.array([[0, 0], [1., 0]]), np.array([[0, 1.], [0, 0]])]
    worst = 0; nbad = 0
    for A in cases:
        u, s, vt = np.linalg.svd(A)
        U, S, VT = svd2(A)
        err = max(np.abs(u - U).max(), np.abs(vt - VT).max(), np.abs(s - S).max() / max(s.max(), 1e-300))
        if err > 1e-9:
            nbad += 1
            if nbad < 6:
                print("MISMATCH", A.tolist(), "\n numpy", u.tolist(), s, vt.tolist(), "\n emul ", U.tolist(), S, VT.tolist())
        else:
            worst = max(worst, err)
    print(len(cases), "cases; mismatches:", nbad, "worst agreeing err", worst)
