"""The algebra DESIGN.md section 1 rests on, checked in numpy against the reference's own formulation (CPU):
DCT-I == Re FFT of the even embedding; the cropped circulant operators are Toeplitz products that any embedding
length L >= 2m-1 (L' >= N+m-1 for R^T / R) reproduces -- including the 1e-6 clamp."""
import numpy as np


def embed_even(c):
    out = c
    for d in range(c.ndim):
        rev = np.flip(out, d); sl = [slice(None)] * c.ndim; sl[d] = slice(1, -1)
        out = np.concatenate([out, rev[tuple(sl)]], d)
    return out


def dct1(x, axis):
    m = x.shape[axis]; j = np.arange(m); k = np.arange(m)
    W = 2 * np.cos(np.pi * np.outer(k, j) / (m - 1)); W[:, 0] = 1; W[:, -1] = (-1.0) ** k
    return np.moveaxis(np.tensordot(W, np.moveaxis(x, axis, 0), 1), 0, axis)


def test_identities():
    rng = np.random.default_rng(0)
    m = (7, 5); N = tuple(2 * a - 2 for a in m)
    g1 = np.linspace(0, 1, m[0]); g2 = np.linspace(0, 2, m[1])
    r = np.sqrt(g1[:, None] ** 2 + g2[None, :] ** 2)
    c = np.exp(-r / 0.2) * (1 + r); c[0, 0] += 1e-3
    Dfull = np.fft.fftn(embed_even(c)).real
    assert np.abs(dct1(dct1(c, 0), 1) - Dfull[:m[0], :m[1]]).max() < 1e-12
    D = np.maximum(Dfull, 0.5)                      # heavy clamping on purpose

    def ref(spec, v, pad=True, crop=True):
        z = np.zeros(N)
        if pad:
            z[:m[0], :m[1]] = v.reshape(m)
        else:
            z = v.reshape(N)
        o = np.fft.ifftn(spec * np.fft.fftn(z)).real
        return o[:m[0], :m[1]].ravel() if crop else o.ravel()

    v = rng.standard_normal(m[0] * m[1]); w = rng.standard_normal(N[0] * N[1])
    idct = lambda Dm: dct1(dct1(Dm, 0), 1) / np.prod(N)
    for f in (lambda d: d, lambda d: 1 / d):
        col = idct(f(D[:m[0], :m[1]]))
        L = (16, 10)
        h = np.zeros(L)
        for i in range(-(m[0] - 1), m[0]):
            for j in range(-(m[1] - 1), m[1]):
                h[i % L[0], j % L[1]] = col[abs(i), abs(j)]
        S = np.fft.fftn(h)
        assert np.abs(S.imag).max() < 1e-12
        z = np.zeros(L); z[:m[0], :m[1]] = v.reshape(m)
        o = np.fft.ifftn(S * np.fft.fftn(z)).real[:m[0], :m[1]].ravel()
        assert np.abs(o - ref(f(D), v)).max() < 1e-12
    s = idct(np.sqrt(D[:m[0], :m[1]]))
    Lp = (18, 12)
    fold = lambda k, d: (k % N[d]) if (k % N[d]) < m[d] else N[d] - (k % N[d])
    h = np.zeros(Lp)
    for i in range(-(m[0] - 1), N[0]):
        for j in range(-(m[1] - 1), N[1]):
            h[i % Lp[0], j % Lp[1]] = s[fold(i, 0), fold(j, 1)]
    S = np.fft.fftn(h)
    z = np.zeros(Lp); z[:m[0], :m[1]] = v.reshape(m)
    o = np.fft.ifftn(S * np.fft.fftn(z)).real[:N[0], :N[1]].ravel()
    assert np.abs(o - ref(np.sqrt(D), v, crop=False)).max() < 1e-12
    z = np.zeros(Lp); z[:N[0], :N[1]] = w.reshape(N)
    o = np.fft.ifftn(np.conj(S) * np.fft.fftn(z)).real[:m[0], :m[1]].ravel()
    assert np.abs(o - ref(np.sqrt(D), w, pad=False)).max() < 1e-12
