"""Randomized Hadamard rotations for the QuaRot / ViDiT-Q layers.

Counterpart of ViDiT-Q/quant_utils/qdiff/quarot/quarot_utils.py (`get_hadK` :100-154, `matmul_hadU` :158-179,
`random_hadamard_matrix` :186-192), built from scratch: the reference ships 97k lines of literal +-1 tables (and loads
its order-144 matrix from a hard-coded absolute path, :261-264); here every base matrix is CONSTRUCTED:

  order 2^m          Sylvester
  order q+1, q prime, q = 3 (mod 4)   Paley construction I  (12 = 11+1, 20 = 19+1, 28 -> see below, 108 = 107+1, 140 = 139+1)
  order 2(q+1), q prime, q = 1 (mod 4)  Paley construction II (28 = 2*(13+1), 36 = 2*(17+1), 52 = 2*(25+1) not prime -> n/a)
  products          Kronecker (40 = 2 x 20, 156 = ... )

`hadamard_factor(n)` picks n = K * 2^m with the smallest constructible K, which also fixes reference defect B-11
(SURVEY appendix B): 13824 (Wan-14B ffn) = 108 * 2^7, where the reference tries 144 * 96 first and asserts.
Wan sizes: 1536 = 12*2^7, 8960 = 140*2^6, 5120 = 20*2^8 (the reference uses 40*2^7), 13824 = 108*2^7.

The rotation a layer applies is R = diag(s) . H_n / sqrt(n), s = random +-1 (random_hadamard_matrix), orthogonal:
(x R)(W R)^T = x W^T.  Any Hadamard matrix of the right order gives a valid rotation; `rotation_matrix` is regenerated
on load in the reference as well (quant_model.py:145-152), so checkpoints stay interchangeable.
"""
import math

import torch


def _is_prime(q):
    if q < 2:
        return False
    for d in range(2, int(math.isqrt(q)) + 1):
        if q % d == 0:
            return False
    return True


def _jacobsthal(q):
    """Q[i,j] = chi(i - j) over GF(q), q prime: chi = quadratic-residue character."""
    res = {(x * x) % q for x in range(1, q)}
    chi = torch.tensor([0] + [1 if x in res else -1 for x in range(1, q)], dtype=torch.float64)
    idx = (torch.arange(q).view(-1, 1) - torch.arange(q).view(1, -1)) % q
    return chi[idx]


def _paley1(q):
    """order q+1, q prime = 3 mod 4:  H = I + S,  S = [[0, 1^T], [-1, Q]] skew-symmetric."""
    n = q + 1
    S = torch.zeros(n, n, dtype=torch.float64)
    S[0, 1:] = 1
    S[1:, 0] = -1
    S[1:, 1:] = _jacobsthal(q)
    return torch.eye(n, dtype=torch.float64) + S


def _paley2(q):
    """order 2(q+1), q prime = 1 mod 4: S symmetric; entries 0 -> [[1,-1],[-1,-1]], +-1 -> +-[[1,1],[1,-1]]."""
    n = q + 1
    S = torch.zeros(n, n, dtype=torch.float64)
    S[0, 1:] = 1
    S[1:, 0] = 1
    S[1:, 1:] = _jacobsthal(q)
    A = torch.tensor([[1., 1.], [1., -1.]], dtype=torch.float64)
    Z = torch.tensor([[1., -1.], [-1., -1.]], dtype=torch.float64)
    H = torch.kron(S, A) + torch.kron(torch.eye(n, dtype=torch.float64), Z)
    return H


_BASE_CACHE = {}


def base_hadamard(K):
    """A Hadamard matrix of order K (float64, entries +-1), or None when no construction here covers K."""
    if K in _BASE_CACHE:
        return _BASE_CACHE[K]
    H = None
    if K == 1:
        H = torch.ones(1, 1, dtype=torch.float64)
    elif K == 2:
        H = torch.tensor([[1., 1.], [1., -1.]], dtype=torch.float64)
    elif K % 4 == 0:
        if _is_prime(K - 1) and (K - 1) % 4 == 3:
            H = _paley1(K - 1)
        elif K % 2 == 0 and _is_prime(K // 2 - 1) and (K // 2 - 1) % 4 == 1:
            H = _paley2(K // 2 - 1)
        else:
            half = base_hadamard(K // 2)
            if half is not None:
                H = torch.kron(base_hadamard(2), half)
    if H is not None:
        assert torch.equal(H @ H.t(), K * torch.eye(K, dtype=torch.float64)), f"order-{K} construction is not Hadamard"
    _BASE_CACHE[K] = H
    return H


def hadamard_factor(n):
    """n = K * 2^m with the smallest K that has a construction (K odd multiple part times a power of two)."""
    m = 0
    while n % 2 == 0 and n > 1:
        n //= 2
        m += 1
    K = n                                       # odd part
    while base_hadamard(K) is None:
        if m == 0:
            raise ValueError("no Hadamard construction for this size")
        K *= 2
        m -= 1
    return K, m


def is_pow2(n):
    return n > 0 and (n & (n - 1)) == 0


def get_hadK(n, transpose=False):
    """(hadK [K,K] or None, K) as quarot_utils.py:100-154."""
    K, _ = hadamard_factor(n)
    if K == 1:
        return None, 1
    H = base_hadamard(K)
    return (H.t() if transpose else H), K


def matmul_hadU(X, transpose=False):
    """X @ (H_n / sqrt(n)) along the last dim with H_n = H_K (x) H_{2^m} (index = a*2^m + b): a natural-order fast
    Walsh-Hadamard transform over the low m index bits, then the order-K base block over the K segments.  Same operator
    as the reference's matmul_hadU (quarot_utils.py:158-179: butterfly stages down to K rows, then hadK @)."""
    n = X.shape[-1]
    K, m = hadamard_factor(n)
    width = 1 << m
    y = X.reshape(-1, K, width).clone()
    h = 1
    while h < width:
        y = y.view(-1, K, width // (2 * h), 2, h)
        lo, hi = y[..., 0, :], y[..., 1, :]
        y = torch.stack((lo + hi, lo - hi), dim=-2)
        h *= 2
    y = y.reshape(-1, K, width)
    if K > 1:
        HK = base_hadamard(K).to(y)
        y = torch.einsum("ij,bjk->bik", HK.t() if transpose else HK, y)
    return y.reshape(X.shape) / math.sqrt(n)


def hadamard_kernel_plan(n):
    """Factorisation b200q_had_quant_rows accepts (include/b200q.h): n = K' * 2^w with w in [5, 8], K' <= 32 and
    n % 128 == 0 -> (K', w, H_K' [K', K'] float64 or None when K' == 1); None when the fused kernel cannot serve n.
    When the power-of-two part of n exceeds 2^8 the surplus Sylvester factor moves into the base block:
    H_n = H_K (x) H_{2^m} = (H_K (x) H_{2^(m-8)}) (x) H_{2^8}."""
    try:
        K, m = hadamard_factor(n)
    except ValueError:
        return None
    if n % 128 != 0 or m < 5:
        return None
    H = base_hadamard(K)
    if m > 8:
        extra = m - 8
        S = torch.ones(1, 1, dtype=torch.float64)
        for _ in range(extra):
            S = torch.kron(base_hadamard(2), S)
        H = torch.kron(H, S)
        K, m = K << extra, 8
    if K > 32:
        return None
    return K, m, (None if K == 1 else H.contiguous())


def matmul_hadUt(X):
    return matmul_hadU(X, transpose=True)


def random_hadamard_matrix(size, device, generator=None):
    """diag(random +-1) . H_size / sqrt(size), float64 [size, size] (quarot_utils.py:186-192)."""
    s = torch.randint(low=0, high=2, size=(size,), generator=generator).to(torch.float64) * 2 - 1
    return matmul_hadU(torch.diag(s)).to(device)
