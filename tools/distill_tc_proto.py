"""CPU prototype of the tensor-core formulation of the teacher-distill kernel (csrc/distill_tc.cu).

Checks the algebra (Bernstein-basis expansion, GEMM form, backward chain) and the split-tf32 error against
the fp64 evaluation of the oracle.  Test infrastructure only."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle import bacs_oracle as O

def conv2(a, b):
    """2-D polynomial product over the two leading axes; trailing axes broadcast."""
    ra, sa = a.shape[:2]; rb, sb = b.shape[:2]
    out = np.zeros((ra + rb - 1, sa + sb - 1) + np.broadcast_shapes(a.shape[2:], b.shape[2:]), dtype=np.result_type(a, b))
    for i in range(ra):
        for j in range(sa):
            for k in range(rb):
                for l in range(sb):
                    out[i + k, j + l] += a[i, j] * b[k, l]
    return out

def tf32_hi(x):
    return (x.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)

def split_mm(A, Bm):
    """A [M,K] x Bm [K,N] with both operands split hi/lo in tf32, three products, fp32 accumulation."""
    A = A.astype(np.float32); Bm = Bm.astype(np.float32)
    Ah = tf32_hi(A); Al = tf32_hi(A - Ah)
    Bh = tf32_hi(Bm); Bl = tf32_hi(Bm - Bh)
    return (Ah @ Bh + Ah @ Bl + Al @ Bh).astype(np.float32)

def run(B, A, h, w, H, W, seed=0, near=None, emulate=True, dtype=np.float64):
    g = torch.Generator().manual_seed(seed)
    old = torch.randn(B, A, h, w, generator=g, dtype=torch.float64)
    new = torch.randn(B, A, h, w, generator=g, dtype=torch.float64) if near is None else old + near * torch.randn(B, A, h, w, generator=g, dtype=torch.float64)
    old = old.float().double(); new = new.float().double()
    m = (torch.rand(B, H, W, generator=g) > 0.4)
    lab = torch.where(m, 0, 1)
    nw = new.clone().requires_grad_(True)
    coef = 1.0
    def up64(x):   # the oracle's up-sample with its fp32 tables, evaluated in fp64 (the oracle itself computes in fp32)
        y0, y1, wy = O._lerp_table(H, h, False); x0, x1, wx = O._lerp_table(W, w, False)
        wy = wy.double().view(-1, 1); wx = wx.double()
        rows = x[..., y0, :] * (1 - wy) + x[..., y1, :] * wy
        return rows[..., x0] * (1 - wx) + rows[..., x1] * wx
    want = torch.linalg.vector_norm((up64(old) ** 2 - up64(nw) ** 2) * m.unsqueeze(1), 2.0, dim=-1).sum()
    want.backward()
    wantg = nw.grad.numpy()

    y0, y1, wy = [t.numpy() for t in O._lerp_table(H, h, False)]
    x0, x1, wx = [t.numpy() for t in O._lerp_table(W, w, False)]
    wy = np.where(y1 == y0, 0.0, wy).astype(np.float64); wx = np.where(x1 == x0, 0.0, wx).astype(np.float64)
    o = old.numpy(); n = new.numpy(); mk = m.numpy().astype(np.float64)
    phi4 = lambda t: np.stack([(1 - t) ** (4 - k) * t ** k for k in range(5)], 0)   # [5, len]
    PX = phi4(wx)            # [5, W]
    PY = phi4(wy)            # [5, H]
    loss = 0.0
    grad = np.zeros_like(n)
    for b in range(B):
        for i in range(h):
            rows = np.nonzero(y0 == i)[0]
            if len(rows) == 0: continue
            i1 = min(i + 1, h - 1)
            j1 = np.minimum(np.arange(w) + 1, w - 1)
            # corner arrays [2(ty),2(tx),A,w]
            ob = o[b]; nb = n[b]
            oc = np.stack([np.stack([ob[:, i, :], ob[:, i, :][:, j1]], 0), np.stack([ob[:, i1, :], ob[:, i1, :][:, j1]], 0)], 0)
            nc = np.stack([np.stack([nb[:, i, :], nb[:, i, :][:, j1]], 0), np.stack([nb[:, i1, :], nb[:, i1, :][:, j1]], 0)], 0)
            p = oc - nc; q = oc + nc
            E = conv2(p, q)                 # [3,3,A,w]
            V = conv2(E, E)                 # [5(ty),5(tx),A,w]
            # moments [5(k), rows, w]
            Mk = np.zeros((5, len(rows), w))
            for X in range(W):
                Mk[:, :, x0[X]] += PX[:, X][:, None] * mk[b, rows, X][None, :]
            Bm = PY[:, rows][:, None, :, None] * Mk[None]      # [5(q),5(k),rows,w]
            Vm = V.transpose(2, 3, 0, 1).reshape(A, w * 25)     # [A, (j,q,k)]
            Bmm = Bm.transpose(3, 0, 1, 2).reshape(w * 25, len(rows))
            S = split_mm(Vm, Bmm).astype(np.float64) if emulate else Vm @ Bmm        # [A, rows]
            rs = np.where(S > 0, 1.0 / np.sqrt(np.maximum(S, 1e-300)), 0.0)
            if emulate: rs = rs.astype(np.float32).astype(np.float64)
            loss += float((S * rs).sum())
            Wm = split_mm(rs, Bmm.T).astype(np.float64) if emulate else rs @ Bmm.T      # [A, (j,q,k)]
            Wt = Wm.reshape(A, w, 5, 5).transpose(2, 3, 0, 1)       # [5,5,A,w]
            if emulate:
                E = E.astype(np.float32).astype(np.float64)
            dE = np.zeros_like(E)
            for r in range(3):
                for s in range(3):
                    for r2 in range(3):
                        for s2 in range(3):
                            dE[r, s] += Wt[r + r2, s + s2] * E[r2, s2]
            # dL/dn = -dp + dq = -2 corr(dE, n)
            dn = np.zeros_like(nc)
            for c in range(2):
                for d in range(2):
                    for r in range(3):
                        for s in range(3):
                            if 0 <= r - c < 2 and 0 <= s - d < 2:
                                dn[c, d] += dE[r, s] * (-2.0 * nc[r - c, s - d])
            # scatter the corner gradients
            for j in range(w):
                grad[b, :, i, j] += dn[0, 0][:, j]; grad[b, :, i, j1[j]] += dn[0, 1][:, j]
                grad[b, :, i1, j] += dn[1, 0][:, j]; grad[b, :, i1, j1[j]] += dn[1, 1][:, j]
    grad *= coef
    rel_loss = abs(loss - float(want.detach())) / abs(float(want.detach())) if float(want.detach()) != 0 else abs(loss)
    rel_grad = np.abs(grad - wantg).max() / max(np.abs(wantg).max(), 1e-300)
    return rel_loss, rel_grad, loss, float(want.detach())

if __name__ == "__main__":
    for emu in (False, True):
        for near in (None, 1e-2, 1e-4):
            print("emulate", emu, "near", near, run(2, 6, 4, 6, 64, 96, near=near, emulate=emu)[:2])
    print("headline-like", run(1, 4, 32, 32, 512, 512, emulate=True)[:2])
    print("headline-like near", run(1, 4, 32, 32, 512, 512, near=1e-3, emulate=True)[:2])
