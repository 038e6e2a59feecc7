# validate the HALF-mode two-stage formulation (math only, complex128)
import numpy as np
def run(N, N1, N2):
    Nc = N // 2; assert N1 * N2 == Nc
    rng = np.random.default_rng(0); x = rng.standard_normal(N)
    z = x[0::2] + 1j * x[1::2]
    # stage 1
    E = np.zeros((N1, N2), complex)  # E[slot][n2]
    def slot(k1): return k1 if k1 <= N1 // 2 else 3 * N1 // 2 - k1
    for n2 in range(N2):
        v = np.array([z[N2 * n1 + n2] for n1 in range(N1)])
        V = np.fft.fft(v)
        for k1 in range(N1):
            E[slot(k1), n2] = V[k1] * np.exp(-2j * np.pi * n2 * k1 / Nc)
    X = np.zeros(Nc + 1, complex); seen = np.zeros(Nc + 1, int)
    def post(Zk, Zm, k):
        a, b = Zk.real, Zk.imag; c, d = Zm.real, Zm.imag
        Ex, Ey = 0.5 * (a + c), 0.5 * (b - d); Ox, Oy = 0.5 * (b + d), 0.5 * (c - a)
        wr, wi = np.cos(2 * np.pi * k / N), -np.sin(2 * np.pi * k / N)
        Tx, Ty = wr * Ox - wi * Oy, wr * Oy + wi * Ox
        X[k] = (Ex + Tx) + 1j * (Ey + Ty); seen[k] += 1
        X[Nc - k] = (Ex - Tx) - 1j * (Ey - Ty); seen[Nc - k] += 1
    for u in range(N1 // 2):
        if u >= 1:
            A = np.fft.fft(E[u]); B = np.fft.fft(E[N1 // 2 + u])
            for k2 in range(N2):
                post(A[k2], B[N2 - 1 - k2], u + N1 * k2)
        else:
            A = np.fft.fft(E[0]); B = np.fft.fft(E[N1 // 2])
            for k2 in range(N2 // 2 + 1):
                post(A[k2], A[(N2 - k2) % N2], N1 * k2)
            for k2 in range((N2 + 1) // 2):
                post(B[k2], B[N2 - 1 - k2], N1 // 2 + N1 * k2)
    ref = np.fft.rfft(x)
    print(N, N1, N2, "max err", np.abs(X - ref).max(), "coverage", seen.min(), seen.max(), np.bincount(seen))
run(400, 20, 10); run(512, 16, 16); run(1024, 32, 16); run(2048, 32, 32)
