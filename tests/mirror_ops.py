"""TEST INFRASTRUCTURE ONLY -- torch (CPU or CUDA) mirror of the C-ABI kernels in include/csi_that.h.

Two uses, both in tests/:
  * not-gpu tests inject it into THATEngine to check the host-side sequencing and the hand-derived backward
    formulas against the oracle without a GPU;
  * gpu tests call it op-by-op, on the same inputs, as the per-kernel expected value.
The product package never imports this file.  Dropout / augmentation are not mirrored (p must be 0).
"""
import math

import torch

LEAKY = 0.01


def _rows_view(a, shift, rows):
    """Rows [shift, shift+rows) relative to the body of a token buffer (reads halo/guard rows)."""
    ld = a.stride(0)
    return a.as_strided((rows, a.shape[1]), (ld, 1), a.storage_offset() + shift * ld)


def _valid_rows(B, L, halo, device):
    Lp = L + 2 * halo
    idx = torch.arange(B * Lp, device=device)
    l = idx % Lp - halo
    return (l >= 0) & (l < L)


class MirrorOps:
    def __init__(self, device="cpu"):
        self.device = torch.device(device)

    def make_pack_table(self, entries, device):
        return list(entries)

    # ------------------------------------------------------------------ pack
    @staticmethod
    def _padded_index(n, grp, device):
        """compact index -> head-padded index (csi_grp)."""
        idx = torch.arange(n, device=device)
        valid, pad = grp
        return idx if not pad else (idx // valid) * pad + idx % valid

    def pack_weights(self, params, packed, table, n_entries, max_elems):
        for (src, dst, N, C, k, ld, mode, P, seg_base, gn, gc) in table:
            w = params[src:src + N * C * k].view(N, C, k)
            ni, ci = self._padded_index(N, gn, params.device), self._padded_index(C, gc, params.device)
            rows = (int(ni.max()) if mode == 0 else int(ci.max())) + 1
            m = packed[dst:dst + rows * ld].view(rows, ld)
            for j in range(k):
                if mode == 0:
                    m[ni[:, None], (j * P + ci)[None, :]] = w[:, :, j].to(packed.dtype)
                else:
                    m[ci[:, None], ((seg_base + j) * P + ni)[None, :]] = w[:, :, j].t().to(packed.dtype)

    # ------------------------------------------------------------------ input stage
    def pool_dual(self, x, offs, lens, B, T, F, pe, left, right, halo, augment, rng):
        assert not augment
        L = T // 20
        if offs is not None:
            dense = torch.zeros(B, T, F, dtype=torch.float32, device=x.device)
            flat = x.reshape(-1)
            for b in range(B):
                t = int(lens[b])
                dense[b, T - t:] = flat[int(offs[b]):int(offs[b]) + t * F].view(t, F)
            x = dense
        pooled = x[:B].view(B, L, 20, F).sum(dim=2) * (1.0 / 20)               # [B, L, F]
        Lp_l, Lp_r = L + 2 * halo, F + 2 * halo
        lv = left[:B * Lp_l].view(B, Lp_l, -1)
        lv[:, halo:halo + L, :F] = pooled + (pe[:, :F] if pe is not None else 0)
        rv = right[:B * Lp_r].view(B, Lp_r, -1)
        rv[:, halo:halo + F, :L] = pooled.transpose(1, 2)

    def gauss_pe_fwd(self, pos, mu, sigma, emb, L, K, F, w, pe):
        pos, mu, sigma, emb = pos.view(L, K), mu.view(1, K), sigma.view(1, K), emb.view(K, F)
        diff = pos - mu
        logp = -(diff * diff) / sigma / sigma / 2 - torch.log(sigma)
        ww = torch.softmax(logp, dim=-1)
        w.view(L, K).copy_(ww)
        pe[:, :F] = ww @ emb

    def gauss_pe_bwd(self, dleft, B, halo, w, pos, mu, sigma, emb, L, K, F, dpe_ws, demb, dmu, dsigma):
        Lp = L + 2 * halo
        dpe = dleft[:B * Lp].view(B, Lp, -1)[:, halo:halo + L, :F].sum(0)          # [L, F]
        dpe_ws[:, :F] = dpe
        pos, mu, sigma, emb, w = pos.view(L, K), mu.view(1, K), sigma.view(1, K), emb.view(K, F), w.view(L, K)
        demb.view(K, F).add_(w.t() @ dpe)
        dw = dpe @ emb.t()
        dlogp = w * (dw - (w * dw).sum(-1, keepdim=True))
        diff = pos - mu
        dmu.view(1, K).add_((dlogp * diff / (sigma * sigma)).sum(0, keepdim=True))
        dsigma.view(1, K).add_((dlogp * (diff * diff / sigma ** 3 - 1 / sigma)).sum(0, keepdim=True))

    # ------------------------------------------------------------------ layernorm
    def layernorm_fwd(self, x, gamma, beta, y, mean, rstd, B, L, d, halo, eps):
        rows = B * (L + 2 * halo)
        v = _valid_rows(B, L, halo, x.device)
        xx = x[:rows, :d].float()
        mu = xx.mean(-1)
        var = xx.var(-1, unbiased=False)
        rs = torch.rsqrt(var + eps)
        yy = (xx - mu[:, None]) * rs[:, None] * gamma + beta
        y[:rows].zero_()
        y[:rows, :d] = torch.where(v[:, None], yy, torch.zeros_like(yy)).to(y.dtype)
        mean[:rows] = torch.where(v, mu, torch.zeros_like(mu))
        rstd[:rows] = torch.where(v, rs, torch.zeros_like(rs))

    def layernorm_bwd(self, dy, x, gamma, mean, rstd, dres, dx, dxm, drop_p, drop_site, rng, dgamma, dbeta,
                      B, L, d, halo):
        assert drop_p == 0.0
        rows = B * (L + 2 * halo)
        v = _valid_rows(B, L, halo, x.device)[:, None]
        dyv = dy[:rows, :d].float()
        xh = (x[:rows, :d] - mean[:rows, None]) * rstd[:rows, None]
        gg = dyv * gamma
        dxv = rstd[:rows, None] * (gg - gg.mean(-1, keepdim=True) - xh * (gg * xh).mean(-1, keepdim=True))
        if dres is not None:
            dxv = dxv + dres[:rows, :d]
        dxv = torch.where(v, dxv, torch.zeros_like(dxv))
        dx[:rows].zero_()
        dx[:rows, :d] = dxv
        if dxm is not None:
            dxm[:rows].zero_()
            dxm[:rows, :d] = dxv.to(dxm.dtype)
        dgamma.add_(torch.where(v, dyv * xh, torch.zeros_like(dyv)).sum(0))
        dbeta.add_(torch.where(v, dyv, torch.zeros_like(dyv)).sum(0))

    # ------------------------------------------------------------------ GEMMs
    def gemm_nt_banded(self, A, Bw, C, M, N, segs, bands, bias, residual, drop_p, drop_site, rng):
        # the weights are zero outside a segment's band, so the plain contraction is the banded one
        for (shift, aoff, boff, klen), (lo, hi) in zip(segs, bands):
            w = Bw[:N, boff:boff + klen].float()
            assert float(w[:lo].abs().sum()) == 0.0 and float(w[hi:].abs().sum()) == 0.0
        self.gemm_nt(A, Bw, C, M, N, segs, bias, residual, drop_p, drop_site, rng)

    def gemm_nt(self, A, Bw, C, M, N, segs, bias, residual, drop_p, drop_site, rng):
        assert drop_p == 0.0
        acc = torch.zeros(M, N, dtype=torch.float32, device=A.device)
        for (shift, aoff, boff, klen) in segs:
            a = _rows_view(A, shift, M)[:, aoff:aoff + klen].float()
            b = Bw[:N, boff:boff + klen].float()
            acc += a @ b.t()
        if bias is not None:
            acc += bias[:N]
        if residual is not None:
            acc += residual[:M, :N]
        C[:M, :N] = acc.to(C.dtype)

    @staticmethod
    def _compact_cols(n, grp, device):
        """indices of the non-padding entries among n head-padded entries."""
        idx = torch.arange(n, device=device)
        valid, pad = grp
        return idx if not pad else idx[idx % pad < valid]

    def gemm_tn(self, A, Bv, C, ldc, c_col_stride, M, Na, segs, i_grp=(0, 0), q_grp=(0, 0)):
        isel = self._compact_cols(Na, i_grp, A.device)
        a = A[:M, :Na].float()[:, isel]
        for (shift, boff, coff, nlen) in segs:
            qsel = self._compact_cols(nlen, q_grp, A.device)
            b = _rows_view(Bv, shift, M)[:, boff:boff + nlen].float()[:, qsel]
            r = a.t() @ b                                               # [Na_compact, nlen_compact]
            cv = C.as_strided((len(isel), len(qsel)), (ldc, c_col_stride), C.storage_offset() + coff)
            cv.add_(r)

    def colsum_tokens(self, A, B, L, halo, ncols, out, grp=(0, 0)):
        rows = B * (L + 2 * halo)
        v = _valid_rows(B, L, halo, A.device)[:, None]
        a = A[:rows, :ncols].float()
        sel = self._compact_cols(ncols, grp, A.device)
        out[:len(sel)].add_(torch.where(v, a, torch.zeros_like(a)).sum(0)[sel])

    # ------------------------------------------------------------------ attention
    @staticmethod
    def _heads(buf, B, L, H, hd, hp, halo, which=0):
        """[B,H,L,hd] view (copy) of head-padded columns [which*H*hp + h*hp, +hd) of a token buffer."""
        Lp = L + 2 * halo
        t = buf[:B * Lp].view(B, Lp, -1)[:, halo:halo + L, which * H * hp:(which + 1) * H * hp].float()
        return t.reshape(B, L, H, hp)[..., :hd].permute(0, 2, 1, 3)

    @staticmethod
    def _put_heads(buf, val, B, L, H, hd, hp, halo, which=0):
        Lp = L + 2 * halo
        dst = buf[:B * Lp].view(B, Lp, -1)[:, halo:halo + L, which * H * hp:(which + 1) * H * hp]
        full = torch.zeros(B, L, H, hp, dtype=torch.float32, device=buf.device)
        full[..., :hd] = val.permute(0, 2, 1, 3)
        dst.copy_(full.reshape(B, L, H * hp).to(buf.dtype))

    def _split(self, qkv, B, L, d, H, hp, halo):
        hd = d // H
        return tuple(self._heads(qkv, B, L, H, hd, hp, halo, w) for w in range(3))

    def attn_fwd(self, qkv, o, lse, B, L, d, H, hp, halo):
        hd = d // H
        q, k, v = self._split(qkv, B, L, d, H, hp, halo)
        s = (q @ k.transpose(-1, -2)) * (1.0 / math.sqrt(hd))
        l = torch.logsumexp(s, dim=-1)
        p = torch.exp(s - l[..., None])
        self._put_heads(o, p @ v, B, L, H, hd, hp, halo)
        lse[:B * H * L] = l.reshape(-1)

    def attn_bwd(self, qkv, o, dout, dqkv, lse, B, L, d, H, hp, halo, dbias=None):
        hd = d // H
        sc = 1.0 / math.sqrt(hd)
        q, k, v = self._split(qkv, B, L, d, H, hp, halo)
        oo, do = self._heads(o, B, L, H, hd, hp, halo), self._heads(dout, B, L, H, hd, hp, halo)
        l = lse[:B * H * L].view(B, H, L)
        p = torch.exp((q @ k.transpose(-1, -2)) * sc - l[..., None])
        dv = p.transpose(-1, -2) @ do
        dp = do @ v.transpose(-1, -2)
        D = (do * oo).sum(-1, keepdim=True)
        ds = p * (dp - D) * sc
        dq = ds @ k
        dk = ds.transpose(-1, -2) @ q
        for w, val in enumerate((dq, dk, dv)):
            self._put_heads(dqkv, val, B, L, H, hd, hp, halo, w)
            if dbias is not None:                                   # in_proj bias gradient: compact channel order
                rounded = val.to(dqkv.dtype).float() if dqkv.dtype != torch.float32 else val
                dbias[w * d:(w + 1) * d] += rounded.sum(dim=(0, 2)).reshape(-1)

    # ------------------------------------------------------------------ batchnorm + activation
    def bn_stats(self, z, B, L, halo, ncols, sums):
        rows = B * (L + 2 * halo)
        v = _valid_rows(B, L, halo, z.device)[:, None]
        zz = torch.where(v, z[:rows, :ncols].double(), torch.zeros(1, dtype=torch.float64, device=z.device))
        sums[:ncols].add_(zz.sum(0))
        sums[ncols:2 * ncols].add_((zz * zz).sum(0))

    def bn_finalize(self, sums, Dp, d, nbr, count, conv_bias, run_mean, run_var, nbt, momentum, eps, mean, invstd):
        nc = nbr * Dp
        m = sums[:nc] / count
        var = (sums[nc:2 * nc] / count - m * m).clamp_min(0)
        mean[:nc] = 0
        invstd[:nc] = 0
        for br in range(nbr):
            mean[br * Dp:br * Dp + d] = m[br * Dp:br * Dp + d].float()
            invstd[br * Dp:br * Dp + d] = torch.rsqrt(var[br * Dp:br * Dp + d] + eps).float()
            mb, vb = m[br * Dp:br * Dp + d].float(), var[br * Dp:br * Dp + d].float()
            run_mean[br].mul_(1 - momentum).add_(momentum * (mb + conv_bias[br]))
            run_var[br].mul_(1 - momentum).add_(momentum * vb * (count / max(count - 1, 1)))
            nbt[br].add_(1)

    def bn_eval_prepare(self, Dp, d, nbr, conv_bias, run_mean, run_var, eps, mean, invstd):
        for br in range(nbr):
            mean[br * Dp:br * Dp + d] = run_mean[br] - conv_bias[br]
            invstd[br * Dp:br * Dp + d] = torch.rsqrt(run_var[br] + eps)

    def _bn_common(self, z, mean, invstd, gamma, beta, B, L, d, halo, nbr):
        rows = B * (L + 2 * halo)
        Dp = mean.numel() // nbr
        zh, y = [], []
        for br in range(nbr):
            zz = z[:rows, br * Dp:br * Dp + d].float()
            h = (zz - mean[br * Dp:br * Dp + d]) * invstd[br * Dp:br * Dp + d]
            zh.append(h)
            y.append(h * gamma[br] + beta[br])
        return rows, Dp, zh, y

    def bn_act_fwd(self, z, mean, invstd, gamma, beta, t_res, out, B, L, d, halo, nbr, p_branch, site_branch,
                   p_out, site_out, rng, masks=None):
        assert p_branch == 0.0 and p_out == 0.0
        rows, Dp, zh, y = self._bn_common(z, mean, invstd, gamma, beta, B, L, d, halo, nbr)
        acc = sum(torch.nn.functional.leaky_relu(u, LEAKY) for u in y) / nbr
        v = _valid_rows(B, L, halo, z.device)[:, None]
        res = acc + t_res[:rows, :d]
        out[:rows].zero_()
        out[:rows, :d] = torch.where(v, res, torch.zeros_like(res))

    def _bn_dy(self, dout, y, rows, d, nbr):
        return [dout[:rows, :d] * (1.0 / nbr) * torch.where(u > 0, torch.ones_like(u), torch.full_like(u, LEAKY))
                for u in y]

    def bn_act_bwd_reduce(self, dout, z, mean, invstd, gamma, beta, B, L, d, halo, nbr, p_branch, site_branch,
                          p_out, site_out, rng, red, masks=None):
        assert p_branch == 0.0 and p_out == 0.0
        rows, Dp, zh, y = self._bn_common(z, mean, invstd, gamma, beta, B, L, d, halo, nbr)
        dy = self._bn_dy(dout, y, rows, d, nbr)
        v = _valid_rows(B, L, halo, z.device)[:, None]
        nc = nbr * Dp
        for br in range(nbr):
            dd = torch.where(v, dy[br], torch.zeros_like(dy[br])).double()
            red[br * Dp:br * Dp + d].add_(dd.sum(0))
            red[nc + br * Dp:nc + br * Dp + d].add_((dd * zh[br].double()).sum(0))

    def bn_act_bwd_dz(self, dout, z, mean, invstd, gamma, beta, red, B, L, d, halo, nbr, p_branch, site_branch,
                      p_out, site_out, rng, dz, dgamma, dbeta, masks=None):
        assert p_branch == 0.0 and p_out == 0.0
        rows, Dp, zh, y = self._bn_common(z, mean, invstd, gamma, beta, B, L, d, halo, nbr)
        dy = self._bn_dy(dout, y, rows, d, nbr)
        v = _valid_rows(B, L, halo, z.device)[:, None]
        nc = nbr * Dp
        n = B * L
        dz[:rows].zero_()
        for br in range(nbr):
            s1 = red[br * Dp:br * Dp + d].float()
            s2 = red[nc + br * Dp:nc + br * Dp + d].float()
            r = gamma[br] * invstd[br * Dp:br * Dp + d] * (dy[br] - s1 / n - zh[br] * s2 / n)
            dz[:rows, br * Dp:br * Dp + d] = torch.where(v, r, torch.zeros_like(r)).to(dz.dtype)
            dgamma[br].add_(s2)
            dbeta[br].add_(s1)

    # ------------------------------------------------------------------ heads
    def _head_mask(self, B, L, halo, N, n0, k0, k1, device):
        Lp = L + 2 * halo
        idx = torch.arange(B * Lp, device=device)
        t = (idx % Lp - halo)[:, None]
        k = torch.where(torch.arange(N, device=device) < n0, torch.tensor(k0, device=device),
                        torch.tensor(k1, device=device))[None, :]
        return (t >= 0) & (t <= L - k)

    def head_reduce_fwd(self, p, B, L, halo, N, n0, k0, k1, feat):
        Lp = L + 2 * halo
        m = self._head_mask(B, L, halo, N, n0, k0, k1, p.device)
        a = torch.nn.functional.leaky_relu(p[:B * Lp, :N].float(), LEAKY)
        a = torch.where(m, a, torch.zeros_like(a))
        feat[:B, :N] = a.view(B, Lp, N).sum(1)

    def head_reduce_bwd(self, dfeat, p, B, L, halo, N, n0, k0, k1, dp):
        Lp = L + 2 * halo
        m = self._head_mask(B, L, halo, N, n0, k0, k1, p.device)
        pp = p[:B * Lp, :N].float()
        g = dfeat[:B, :N].float()[:, None, :].expand(B, Lp, N).reshape(B * Lp, N)
        g = g * torch.where(pp > 0, torch.ones_like(pp), torch.full_like(pp, LEAKY))
        dp[:B * Lp].zero_()
        dp[:B * Lp, :N] = torch.where(m, g, torch.zeros_like(g)).to(dp.dtype)

    def dropout_rows(self, inp, out, rows, cols, p, site, rng):
        assert p == 0.0
        out[:rows, :cols] = inp[:rows, :cols].to(out.dtype)

    # ------------------------------------------------------------------ loss / optimizer
    def bce_logits(self, z, y, rows, cols, pos_weight, grad_scale, loss, dz):
        zz, yy = z[:rows, :cols].float(), y[:rows, :cols].float()
        ls = torch.nn.functional.logsigmoid
        loss[0] = -(pos_weight * yy * ls(zz) + (1 - yy) * ls(-zz)).mean()
        if dz is not None:
            sg = torch.sigmoid(zz)
            dz[:rows, :cols] = (sg * (pos_weight * yy + 1 - yy) - pos_weight * yy) * (grad_scale / (rows * cols))

    def smooth_l1(self, z, y, rows, cols, beta, grad_scale, loss, dz):
        d = z[:rows, :cols].float() - y[:rows, :cols].float()
        quad = d.abs() < beta
        loss[0] = torch.where(quad, 0.5 * d * d / beta, d.abs() - 0.5 * beta).mean()
        if dz is not None:
            dz[:rows, :cols] = torch.where(quad, d / beta, torch.sign(d)) * (grad_scale / (rows * cols))

    def perm_ce(self, z, y, B, heads, classes, cpitch, grad_scale, loss, dz, best_perm=None):
        """PermutationMatchingLoss (model/that_multi_head.py:309-342): first minimum over itertools.permutations."""
        from itertools import permutations
        zz = z[:B, :heads * cpitch].float().view(B, heads, cpitch)[:, :, :classes]
        cls = y[:B, :heads * classes].float().view(B, heads, classes).argmax(-1)
        logp = torch.log_softmax(zz, -1)
        cost = -logp.gather(2, cls.unsqueeze(1).expand(B, heads, heads))            # [B, head, slot]
        perms = torch.tensor(list(permutations(range(heads))), dtype=torch.long, device=z.device)
        slot = torch.arange(heads, device=z.device)
        best = perms[(cost[:, perms, slot].sum(-1) / heads).argmin(1)]               # [B, slot] -> head
        loss[0] = cost[torch.arange(B, device=z.device).unsqueeze(1), best, slot].mean()
        if best_perm is not None:
            best_perm[:B * heads] = best.reshape(-1).to(best_perm.dtype)
        if dz is not None:
            g = torch.softmax(zz, -1)
            onehot = torch.zeros_like(g)
            onehot[torch.arange(B, device=z.device).unsqueeze(1), best, cls] = 1.0
            dz[:B, :heads * cpitch] = 0
            dz[:B, :heads * cpitch].view(B, heads, cpitch)[:, :, :classes] = (g - onehot) * (grad_scale / (B * heads))

    def adam_flat(self, p, g, m, v, n, lr, b1, b2, eps, wd, step, grad_scale):
        t = int(step.item())
        gg = g[:n] * grad_scale + wd * p[:n]
        m[:n].mul_(b1).add_(gg, alpha=1 - b1)
        v[:n].mul_(b2).addcmul_(gg, gg, value=1 - b2)
        bc1, bc2 = 1 - b1 ** t, 1 - b2 ** t
        denom = (v[:n].sqrt() / math.sqrt(bc2)).add_(eps)
        p[:n].addcdiv_(m[:n], denom, value=-lr / bc1)

    def advance_counters(self, rng, step):
        if rng is not None:
            rng[1] += 1
        if step is not None:
            step[0] += 1

    def fill_f32(self, t, v):
        t.fill_(v)

    def fill_f64(self, t, v=0.0):
        t.fill_(v)

    def copy_f32(self, dst, src, n):
        dst.reshape(-1)[:n].copy_(src.reshape(-1)[:n])
