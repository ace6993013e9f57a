"""numpy prototype of the 'quad' blind-rotation transform (4 warps per PBS, 64 threads x 16 points per polynomial):
negacyclic FFT-2048 as a folded FFT-1024 = 16 (registers) x 16 (registers, after one exchange) x 4 (registers, after a
second exchange inside groups of 4 lanes), the Fourier-domain product, and the inverse as the TRANSPOSED forward
algorithm in the swapped (re <-> im) domain.  Validates the index algebra against a direct negacyclic product."""
import numpy as np

N, M = 2048, 1024
rng = np.random.default_rng(0)
zeta = np.exp(1j * np.pi / N)
W = lambda n: np.exp(-2j * np.pi / n)

D = np.exp(1j * np.pi * np.arange(16) / 32)                       # D[n1] = zeta^(64 n1)
T = np.exp(1j * np.pi * np.outer(1 - 4 * np.arange(16), np.arange(64)) / N)   # T[k1][n2] = zeta^n2 W1024^(n2 k1)
F16 = W(16) ** np.outer(np.arange(16), np.arange(16))
F4 = W(4) ** np.outer(np.arange(4), np.arange(4))
w64 = W(64) ** np.outer(np.arange(4), np.arange(16))              # w64[q][j]


def forward(c):
    """c[n] complex, n = 64 n1 + n2 -> spectrum S[tau][jj][r], tau = 4 k1 + s  <->  k = k1 + 16 (s + 4 jj + 16 r)"""
    a = c.reshape(16, 64) * D[:, None]                            # a[n1][n2]  (thread n2, regs n1)
    A = F16 @ a                                                   # A[k1][n2]
    A = A * T
    # exchange 1: thread (k1, q) regs m: b = A[k1][4m + q]
    b = A.reshape(16, 16, 4).transpose(0, 2, 1)                   # b[k1][q][m]
    C = b @ F16.T                                                 # C[k1][q][j] = sum_m b W16^(m j)
    C = C * w64[None, :, :]
    # exchange 2: thread (k1, s) regs (jj, q): j = s + 4 jj
    Cs = C.reshape(16, 4, 4, 4)                                   # [k1][q][jj][s]
    Cs = Cs.transpose(0, 3, 2, 1)                                 # [k1][s][jj][q]
    Y = Cs @ F4.T                                                 # [k1][s][jj][r] = sum_q C W4^(q r)
    return Y.reshape(64, 4, 4)


def spectrum_index():
    k1, s, jj, r = np.meshgrid(np.arange(16), np.arange(4), np.arange(4), np.arange(4), indexing="ij")
    return (k1 + 16 * (s + 4 * jj + 16 * r)).reshape(64, 4, 4)


def transposed(win):
    """swapped-domain inverse: out[n] = sum_k in[k] Phi[k][n] with the forward kernel Phi (no conjugation, no 1/M)"""
    Y = win.reshape(16, 4, 4, 4)                                  # [k1][s][jj][r]
    U = Y @ F4                                                    # [k1][s][jj][q] = sum_r in W4^(q r)
    U = U.transpose(0, 3, 2, 1).reshape(16, 4, 16)                # [k1][q][j = s + 4 jj] (exchange 2')
    U = U * w64[None, :, :]
    V = U @ F16                                                   # [k1][q][m] = sum_j U W16^(m j)
    V = V.transpose(0, 2, 1).reshape(16, 64)                      # [k1][n2 = 4 m + q]
    V = V * T
    out = F16.T @ V                                               # exchange 1' then DFT16 over k1: [n1][n2]
    out = out * D[:, None]
    return out.reshape(M)


def swap(z):
    return z.imag + 1j * z.real


def negacyclic_mul(p, g):
    full = np.convolve(p, g)
    res = full[:N].copy()
    res[: N - 1] -= full[N:]
    return res


# 1. forward == direct evaluation X_k = sum_n c_n zeta^n W^(n k)
p = rng.integers(-2**22, 2**22, N).astype(np.float64)
c = p[:M] + 1j * p[M:]
S = forward(c)
k = spectrum_index()
n = np.arange(M)
X = np.array([np.sum(c * zeta**n * W(M) ** (n * kk)) for kk in range(M)])
assert np.abs(S - X[k]).max() < 1e-9 * np.abs(X).max(), np.abs(S - X[k]).max()

# 2. transposed(swap(.)) inverts forward up to the factor M
back = swap(transposed(swap(S))) / M
assert np.allclose(back, c, atol=1e-6), np.abs(back - c).max()

# 3. full external-product style check: ifft(fft(p) * fft(g)) == negacyclic p * g (g small integers)
g = rng.integers(-8, 8, N).astype(np.float64)
Gs = forward(g[:M] + 1j * g[M:])
prod = swap(transposed(swap(S * Gs))) / M
want = negacyclic_mul(p, g)
got = np.concatenate([prod.real, prod.imag])
assert np.allclose(got, want, atol=1e-2), np.abs(got - want).max()
print("quad transform prototype ok: forward, transposed inverse and negacyclic product agree")
