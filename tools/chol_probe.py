"""Timing / accuracy probe for the SparseGPT prologue (A9): reference order (potrf, potri, potrf upper) against the
reversal identity  U = (J chol(J H J) J)^-1  (one factorisation + one triangular inverse)."""
import torch, time, sys
dev = torch.device("cuda", 0)
torch.backends.cuda.matmul.allow_tf32 = False


def make_h(C, T, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    X = torch.randn(T, C, device=dev, generator=g) * (torch.rand(C, device=dev, generator=g) + 0.1)
    H = (2.0 / T) * (X.T @ X)
    return H


def ref_order(H):
    L = torch.linalg.cholesky_ex(H)[0]
    Hi = torch.cholesky_inverse(L)
    return torch.linalg.cholesky_ex(Hi, upper=True)[0]


def flip(H):
    Lf = torch.linalg.cholesky_ex(H.flip(0, 1))[0]          # J H J = Lf Lf^T
    V = Lf.flip(0, 1)                                        # H = V V^T, V upper
    I = torch.eye(H.shape[0], device=dev)
    return torch.linalg.solve_triangular(V, I, upper=True)   # U = V^-1: H^-1 = U^T U


def timeit(f, *a, n=3):
    f(*a); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = f(*a); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts), r


for C, T in ((1408, 32896), (2048, 4096), (5120, 4096), (6144, 32896), (768, 3152), (3072, 3152)):
    H = make_h(C, T)
    H = H + 0.01 * H.diag().mean() * torch.eye(C, device=dev)
    t_potrf, L = timeit(lambda h: torch.linalg.cholesky_ex(h)[0], H)
    t_potri, Hi = timeit(torch.cholesky_inverse, L)
    t_potrf2, U = timeit(lambda h: torch.linalg.cholesky_ex(h, upper=True)[0], Hi)
    t_ref, Ua = timeit(ref_order, H)
    t_flip, Ub = timeit(flip, H)
    H64 = H.double()
    U64 = torch.linalg.cholesky(torch.linalg.inv(H64), upper=True)
    ea = ((Ua.double() - U64).norm() / U64.norm()).item()
    eb = ((Ub.double() - U64).norm() / U64.norm()).item()
    eab = ((Ua - Ub).norm() / Ua.norm()).item()
    print(f"C={C}: potrf {t_potrf:.2f} potri {t_potri:.2f} potrf_upper {t_potrf2:.2f} | reference order {t_ref:.2f} ms, reversal {t_flip:.2f} ms"
          f" | rel err vs fp64: ref {ea:.2e} reversal {eb:.2e}, between {eab:.2e}")
