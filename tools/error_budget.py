"""bf16-mode error budget of the GAViKO hot path, site by site (CPU emulation; ANALYSIS TOOL, not on the product path).

The CUDA bf16 mode keeps residual streams, LayerNorm statistics, softmax and the rank-r side paths in fp32 and rounds to bf16 only where a
tensor-core operand or a saved activation is formed.  This tool restates `engine.GavikoEngine.forward/backward` in torch
fp32 on the CPU with every one of those rounding sites as a switch, so the contribution of each site to the gradient error against the
fp32 golden vectors (tests/golden/*.npz, from the live reference) can be measured one at a time:

    python tools/error_budget.py [case ...] > profiles/error_budget_rNN.txt

For every case and loss it prints  (a) all sites on  = the emulated bf16 mode (compare with the measured GPU figure in profiles/parity_*),
(b) each site alone on,  (c) all sites on but one.  Rounding sites (engine.py line numbers of the round-2 code in brackets):

  forward   patch     bf16 patches and Conv3d weight            [patch_gather -> gemm]
            ln1/ln2   bf16 LayerNorm outputs (GEMM A operands)
            wqkv/wo/w1/w2  bf16 copies of the frozen weights (used by the forward GEMM and its dgrad)
            qkv_out   bf16 q, k, v
            attn_p    bf16 probabilities feeding P V (unnormalised, relative to the running row maximum)
            o_out     bf16 attention output
            act_out   bf16 GELU output
  saved     gelu_grad bf16 saved GELU derivative
  backward  dG_lp     bf16 copy of dG entering the fc2 dgrad GEMM
            dA_out    bf16 (dG W2) * gelu'
            dGm_lp    bf16 copy of d g_mid entering the out-projection dgrad
            dO_out    bf16 dO
            attn_bwd  bf16 P^T and dS feeding the dV / dQ / dK MMAs
            dqkv_out  bf16 dq, dk, dv
  side      tf32      tf32 operands of the rank-r products (down / up projections, window attention, their weight gradients)
"""
import math
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle import gaviko_oracle as O  # noqa: E402
from oracle.cases import GAVIKO_CASES  # noqa: E402
from oracle.golden_fill import golden_labels, golden_volume  # noqa: E402

SITES = ['patch', 'ln1', 'wqkv', 'qkv_out', 'attn_p', 'o_out', 'wo', 'ln2', 'w1', 'act_out', 'w2', 'gelu_grad', 'dG_lp', 'dA_out', 'dGm_lp', 'dO_out',
         'attn_bwd', 'dqkv_out', 'tf32']
ACTIVE = set()


def bf(x):
    return x.bfloat16().float()


def tf32(x):
    return ((x.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


def r(x, site, fn=bf):
    return fn(x) if site in ACTIVE else x


class RoundFwd(torch.autograd.Function):
    """round in forward, straight-through gradient"""
    @staticmethod
    def forward(ctx, x, site):
        return r(x, site)

    @staticmethod
    def backward(ctx, g):
        return g, None


class RoundBwd(torch.autograd.Function):
    """identity in forward, rounds the gradient flowing back"""
    @staticmethod
    def forward(ctx, x, site):
        ctx.site = site
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return r(g, ctx.site), None


class MM(torch.autograd.Function):
    """y = ra(a) @ rb(b)^T;  da = rg(dy) @ rb(b);  db = rg(dy)^T @ ra(a) — a GEMM whose operands are rounded at the given sites."""
    @staticmethod
    def forward(ctx, a, b, sa, sb, sg, fn):
        a_r, b_r = r(a, sa, fn), r(b, sb, fn)
        ctx.save_for_backward(a_r, b_r)
        ctx.sg, ctx.fn = sg, fn
        return a_r @ b_r.transpose(-1, -2)

    @staticmethod
    def backward(ctx, dy):
        a_r, b_r = ctx.saved_tensors
        g = r(dy, ctx.sg, ctx.fn)
        da = g @ b_r if ctx.needs_input_grad[0] else None
        db = None
        if ctx.needs_input_grad[1]:
            if b_r.dim() > 2:
                db = g.transpose(-1, -2) @ a_r
            else:
                db = g.reshape(-1, g.shape[-1]).t() @ a_r.reshape(-1, a_r.shape[-1])
        return da, db, None, None, None, None


def mm(a, b, sa=None, sb=None, sg=None, fn=bf):
    return MM.apply(a, b, sa, sb, sg, fn)


def mm_side(a, b):
    """rank-r side product a @ b^T with tf32 operands in forward and backward"""
    return MM.apply(a, b, 'tf32', 'tf32', 'tf32', tf32)


class GeluSite(torch.autograd.Function):
    """act = bf16(gelu(pre)); backward: dpre = bf16(dact * bf16(gelu'(pre)))  (fc1 epilogue saves the derivative; fc2 dgrad epilogue multiplies)"""
    @staticmethod
    def forward(ctx, pre):
        cdf = 0.5 * (1.0 + torch.erf(pre * 0.7071067811865476))
        grad = cdf + pre * 0.3989422804014327 * torch.exp(-0.5 * pre * pre)
        ctx.save_for_backward(r(grad, 'gelu_grad'))
        return r(pre * cdf, 'act_out')

    @staticmethod
    def backward(ctx, dact):
        (grad,) = ctx.saved_tensors
        return r(dact * grad, 'dA_out')


class FlashAttn(torch.autograd.Function):
    """softmax(q k^T scale) v with the roundings of the tcgen05 kernels: P (unnormalised, max-relative) bf16 for P V, row sum from the unrounded
    fp32 values; backward recomputes P from lse, dS = P (dP - delta) and rounds P^T / dS to bf16 for the accumulating MMAs."""
    @staticmethod
    def forward(ctx, q, k, v, scale):
        s = (q @ k.transpose(-1, -2)) * scale
        m = s.amax(-1, keepdim=True)
        p = torch.exp(s - m)
        l = p.sum(-1, keepdim=True)
        o = r(p, 'attn_p') @ v / l
        o = r(o, 'o_out')
        ctx.save_for_backward(q, k, v, o, m + torch.log(l))
        ctx.scale = scale
        return o

    @staticmethod
    def backward(ctx, do):
        q, k, v, o, lse = ctx.saved_tensors
        do = r(do, 'dO_out')
        s = (q @ k.transpose(-1, -2)) * ctx.scale
        p = torch.exp(s - lse)
        delta = (do * o).sum(-1, keepdim=True)
        dp = do @ v.transpose(-1, -2)
        ds = p * (dp - delta)
        p_r, ds_r = r(p, 'attn_bwd'), r(ds, 'attn_bwd')
        dv = p_r.transpose(-1, -2) @ do
        dq = (ds_r @ k) * ctx.scale
        dk = (ds_r.transpose(-1, -2) @ q) * ctx.scale
        return r(dq, 'dqkv_out'), r(dk, 'dqkv_out'), r(dv, 'dqkv_out'), None


def emulated_forward(sd, img, *, backbone, num_prompts, frame_patch_size, image_patch_size, local_k, DHW, share_factor=1, dim_head=64):
    """engine.GavikoEngine.forward / backward restated with switchable rounding sites (autograd supplies the backward structure)."""
    depth, heads, dim, _ = O.mapping_vit(backbone)
    P = num_prompts
    w = sd['conv_proj.0.weight']
    e = mm(O.patchify(img, frame_patch_size, image_patch_size), w.reshape(w.shape[0], -1), 'patch', 'patch') + sd['conv_proj.0.bias']
    B = e.shape[0]
    pos = sd['pos_embedding']
    g = torch.cat([(sd['prompt_embeddings'] + sd['prompt_positional_embedding']).expand(B, -1, -1), (sd['cls_token'] + pos[:, :1]).expand(B, -1, -1),
                   e + pos[:, 1:]], dim=1)
    loc = e + pos[:, 1:]
    allow = O.window_allow(tuple(DHW), tuple(local_k))
    for i in range(depth):
        s_ = i // share_factor
        # ---- local branch: rank-r products on tf32 operands
        p = f'transformer.local_attns.{s_}.'
        z = mm_side(O.layer_norm(loc, sd[p + 'norm.weight'], sd[p + 'norm.bias']), sd[p + 'proj_down.weight']) + sd[p + 'proj_down.bias']
        q, k, v = mm_side(z, sd[p + 'qkv.weight']).chunk(3, dim=-1)
        sc = mm_side(q, k) * dim ** -0.5
        a = torch.softmax(sc.masked_fill(~allow, float('-inf')), dim=-1)
        ctx_l = mm_side(a, v.transpose(-1, -2))
        loc = mm_side(ctx_l, sd[p + 'proj_up.weight']) + sd[p + 'proj_up.bias'] + loc
        # ---- frozen MHSA
        a_ = f'transformer.attns.{i}.'
        T = g.shape[1]
        h1 = O.layer_norm(g, sd[a_ + 'norm.weight'], sd[a_ + 'norm.bias'])
        qkv = RoundFwd.apply(mm(h1, sd[a_ + 'to_qkv.weight'], 'ln1', 'wqkv', None), 'qkv_out')        # dqkv arrives bf16 from the attention backward
        q, k, v = (t.reshape(B, T, heads, dim_head).transpose(1, 2) for t in qkv.chunk(3, dim=-1))
        o = FlashAttn.apply(q, k, v, dim_head ** -0.5).transpose(1, 2).reshape(B, T, heads * dim_head)
        g = mm(o, sd[a_ + 'to_out.0.weight'], None, 'wo', 'dGm_lp') + sd[a_ + 'to_out.0.bias'] + g
        # ---- Awakening_Prompt: down-projections on tf32 operands, gates / cross-attention exact fp32, up-projection hi/lo split (~exact)
        pp = f'transformer.prompt_projs.{s_}.'
        prompt = awakening_prompt_emul(g, loc, sd, pp, P)
        # ---- frozen MLP
        f_ = f'transformer.mlps.{i}.'
        h2 = O.layer_norm(g, sd[f_ + 'net.0.weight'], sd[f_ + 'net.0.bias'])
        act = GeluSite.apply(mm(h2, sd[f_ + 'net.1.weight'], 'ln2', 'w1', None) + sd[f_ + 'net.1.bias'])
        g = mm(act, sd[f_ + 'net.4.weight'], None, 'w2', 'dG_lp') + sd[f_ + 'net.4.bias'] + g + prompt
    out = O.layer_norm(g, sd['transformer.norm.weight'], sd['transformer.norm.bias'])
    return out[:, :P + 1].mean(dim=1) @ sd['mlp_head.head.weight'].t() + sd['mlp_head.head.bias']


def awakening_prompt_emul(g, loc, sd, p, P):
    wd, bd = sd[p + 'proj_down.0.weight'], sd[p + 'proj_down.0.bias']
    xl = O.quick_gelu(mm_side(g, wd) + bd)
    ll = O.quick_gelu(mm_side(loc, wd) + bd)
    rr = xl.shape[-1]
    pl, cl, il = xl[:, :P], xl[:, P:P + 1], xl[:, P + 1:]
    ca = p + 'cls_analyzer.cls_analyzer_.'
    h = O.layer_norm(cl, sd[ca + '0.weight'], sd[ca + '0.bias'])
    h = F.gelu(h @ sd[ca + '1.weight'].t() + sd[ca + '1.bias'])
    imp = torch.sigmoid(h @ sd[ca + '3.weight'].t() + sd[ca + '3.bias'])
    gb = p + 'gl_balancer.gl_balancer_.'
    gw = torch.sigmoid(O.layer_norm(cl, sd[gb + '0.weight'], sd[gb + '0.bias']) @ sd[gb + '1.weight'].t() + sd[gb + '1.bias'])

    def xattn(tok, qp):
        q = pl @ sd[qp + 'weight'].t() + sd[qp + 'bias']
        a = torch.softmax(torch.einsum('bpd,bnd->bpn', q, tok) * rr ** -0.5, dim=-1)
        return torch.einsum('bpn,bnd->bpd', a, tok)

    ctx_g = xattn(il[:, P + 1:], p + 'global_attention.query_proj.')
    ctx_l = xattn(ll, p + 'local_attention.query_proj.')
    enh = (gw * ctx_g + (1 - gw) * ctx_l) * imp.transpose(1, 2)
    comb = torch.cat([enh, cl, il], dim=1)
    # forward: K-extension of the fc2 GEMM with hi / lo bf16 pairs (~2^-16); backward: d comb = dG Wu and dWu on tf32 operands
    return MM.apply(comb, sd[p + 'proj_up.weight'], None, None, 'tf32', tf32) + sd[p + 'proj_up.bias']


def run(sd0, trainable, img, y, kw, loss_name, golden):
    sd = {k: v.clone() for k, v in sd0.items()}
    for n in trainable:
        sd[n].requires_grad_(True)
    logits = emulated_forward(sd, img, backbone=kw['backbone'], num_prompts=kw['num_prompts'], frame_patch_size=kw['frame_patch_size'],
                              image_patch_size=kw['image_patch_size'], local_k=kw['local_k'], DHW=kw['DHW'], share_factor=kw['share_factor'])
    loss = O.focal_loss(logits, y) if loss_name == 'focal' else O.cross_entropy(logits, y)
    loss.backward()
    num = den = 0.0
    per = []
    for n in trainable:
        ref = golden[f'grad_{loss_name}/{n}'].astype(np.float64)
        d = float(np.linalg.norm(sd[n].grad.double().numpy() - ref))
        rn = float(np.linalg.norm(ref))
        num += d * d
        den += rn * rn
        per.append((d, rn, n))
    lg = golden['logits'].astype(np.float64)
    return float(np.linalg.norm(logits.detach().double().numpy() - lg) / np.linalg.norm(lg)), (num / den) ** 0.5, per, den ** 0.5


def main():
    from helpers import load_golden, sd_from_golden
    global ACTIVE
    torch.set_num_threads(os.cpu_count())
    cases = [a for a in sys.argv[1:] if not a.startswith('-')] or ['gaviko_t16_small', 'gaviko_t16_full']
    quick = '--quick' in sys.argv
    for name in cases:
        kw, batch = GAVIKO_CASES[name]
        g = load_golden(name)
        sd0 = sd_from_golden(g, requires_grad=False)
        trainable = g['trainable_names'].tolist()
        img = golden_volume(batch, kw['frames'], kw['image_size'], kw['image_size'], channels=kw['channels'])
        y = golden_labels(batch, kw['num_classes'])
        for loss_name in ('focal', 'ce'):
            print(f'== {name} | {loss_name} ==   (logits rel-L2, gradients global rel-L2 against the fp32 golden vectors)')
            ACTIVE = set()
            rl, rg, _, _ = run(sd0, trainable, img, y, kw, loss_name, g)
            print(f'  {"no site (fp32 emulation)":34s} logits {rl:.3e}  grads {rg:.3e}')
            ACTIVE = set(SITES)
            rl, rg, per, gn = run(sd0, trainable, img, y, kw, loss_name, g)
            print(f'  {"ALL sites (emulated bf16 mode)":34s} logits {rl:.3e}  grads {rg:.3e}')
            worst = sorted(((d / rn if rn > 0 else 0.0, rn / gn, n) for d, rn, n in per if rn > 1e-3 * gn), reverse=True)[:4]
            for rel, share, n in worst:
                print(f'      worst tensors (> 1e-3 of global): {rel:.3e}  share {share:.1e}  {n}')
            groups = {'weights (wqkv wo w1 w2 patch)': {'wqkv', 'wo', 'w1', 'w2', 'patch'}, 'forward activations': {'ln1', 'ln2', 'qkv_out', 'attn_p', 'o_out', 'act_out'},
                      'backward activations': {'gelu_grad', 'dG_lp', 'dA_out', 'dGm_lp', 'dO_out', 'attn_bwd', 'dqkv_out'}, 'tf32 side paths': {'tf32'}}
            for gname, gs in groups.items():
                ACTIVE = set(gs)
                rl, rg, _, _ = run(sd0, trainable, img, y, kw, loss_name, g)
                print(f'  only {gname:29s} logits {rl:.3e}  grads {rg:.3e}')
            if quick:
                continue
            for s in SITES:
                ACTIVE = {s}
                rl1, rg1, _, _ = run(sd0, trainable, img, y, kw, loss_name, g)
                ACTIVE = set(SITES) - {s}
                rl2, rg2, _, _ = run(sd0, trainable, img, y, kw, loss_name, g)
                print(f'  site {s:10s} alone: logits {rl1:.3e} grads {rg1:.3e}   | all but it: logits {rl2:.3e} grads {rg2:.3e}')


if __name__ == '__main__':
    main()
