"""Python mirror of the planned K2 row kernels: two real rows per complex FFT, Stockham radix-8
(+ one leading radix-2/4 stage), in place with a read phase and a write phase per stage."""
import numpy as np
from scipy.fft import dct, idct

def stages(n):
    L = n.bit_length() - 1
    lead = [0, 2, 4][L % 3]
    return lead, L // 3

def fft_inplace(buf, n):
    """forward FFT of buf[0:n] (complex128) exactly as the kernel does it."""
    tw = np.exp(-2j * np.pi * np.arange(n) / n)
    lead, n8 = stages(n)
    ns = 1
    if lead == 2:
        half = n // 2
        a = buf[:half].copy(); b = buf[half:].copy()
        out = np.empty(n, complex)
        out[0::2] = a + b; out[1::2] = a - b     # j0 = 2j, ns = 1
        buf[:] = out; ns = 2
    elif lead == 4:
        q = n // 4
        v = [buf[i*q:(i+1)*q].copy() for i in range(4)]
        out = np.empty(n, complex)
        y0 = v[0] + v[2]; y1 = v[0] - v[2]; y2 = v[1] + v[3]; y3 = -1j * (v[1] - v[3])
        out[0::4] = y0 + y2; out[1::4] = y1 + y3; out[2::4] = y0 - y2; out[3::4] = y1 - y3
        buf[:] = out; ns = 4
    e = n // 8
    r = np.sqrt(0.5)
    for _ in range(n8):
        tstep = e // ns
        out = np.empty(n, complex)
        for j in range(e):
            k = j & (ns - 1)
            w1 = tw[k * tstep]; w2 = tw[2 * k * tstep]; w4 = tw[4 * k * tstep]
            w3 = w1 * w2; w5 = w1 * w4; w6 = w2 * w4; w7 = w3 * w4
            u = [buf[j + q * e] for q in range(8)]
            ws = [1, w1, w2, w3, w4, w5, w6, w7]
            if ns > 1:
                u = [u[q] * ws[q] for q in range(8)]
            a = [u[i] + u[i + 4] for i in range(4)]
            b = [u[i] - u[i + 4] for i in range(4)]
            b[1] = b[1] * complex(r, -r); b[2] = b[2] * (-1j); b[3] = b[3] * complex(-r, -r)
            def fft4(x):
                t0 = x[0] + x[2]; t1 = x[0] - x[2]; t2 = x[1] + x[3]; t3 = -1j * (x[1] - x[3])
                return [t0 + t2, t1 + t3, t0 - t2, t1 - t3]
            A = fft4(a); B = fft4(b)
            X = [A[0], B[0], A[1], B[1], A[2], B[2], A[3], B[3]]
            j0 = ((j - k) << 3) + k
            for q in range(8):
                out[j0 + q * ns] = X[q]
        buf[:] = out
        ns *= 8
    return buf

def dct2_pair(xa, xb):
    n = len(xa)
    j = np.arange(n)
    dst = np.where(j & 1, n - 1 - (j >> 1), j >> 1)
    buf = np.empty(n, complex)
    buf[dst] = xa + 1j * xb
    fft_inplace(buf, n)
    k = np.arange(n)
    w = np.exp(-1j * np.pi * k / (2 * n))
    Zk = buf; Znk = buf[(n - k) % n]
    Sx = Zk.real + Znk.real; Sy = Zk.imag - Znk.imag
    Dx = Zk.real - Znk.real; Dy = Zk.imag + Znk.imag
    ya = w.real * Sx - w.imag * Sy
    yb = w.real * Dy + w.imag * Dx
    return ya, yb

def idct2_pair(ya, yb):
    n = len(ya)
    k = np.arange(n)
    w = np.exp(-1j * np.pi * k / (2 * n))
    def V(y):
        ynk = np.where(k > 0, y[(n - k) % n], 0.0)
        re = 0.5 * (w.real * y - w.imag * ynk)
        im = 0.5 * (-w.imag * y - w.real * ynk)
        return re, im
    ra, ia = V(ya); rb, ib = V(yb)
    buf = (ra - ib) + 1j * (-ia - rb)
    fft_inplace(buf, n)
    j = np.arange(n)
    src = np.where(j & 1, n - 1 - (j >> 1), j >> 1)
    return buf.real[src] / n, -buf.imag[src] / n

if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for n in [2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048]:
        z = rng.normal(size=n) + 1j * rng.normal(size=n)
        e1 = np.abs(fft_inplace(z.copy(), n) - np.fft.fft(z)).max()
        xa, xb = rng.normal(size=n), rng.normal(size=n)
        ya, yb = dct2_pair(xa, xb)
        e2 = max(np.abs(ya - dct(xa)).max(), np.abs(yb - dct(xb)).max())
        za, zb = idct2_pair(ya, yb)
        e3 = max(np.abs(za - xa).max(), np.abs(zb - xb).max())
        print(n, stages(n), f"{e1:.1e} {e2:.1e} {e3:.1e}")
