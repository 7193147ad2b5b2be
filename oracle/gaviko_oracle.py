"""fp32 torch restatement of the reference hot path (TEST INFRASTRUCTURE ONLY).

Every function is a closed-form restatement of the cited reference lines
(paths relative to ``/root/reference/src``), written against a plain
``state_dict`` so it needs neither the reference nor ``gaviko_b200``.  It is
pinned against outputs of the live reference by ``oracle/make_golden.py`` →
``tests/golden/*.npz`` (``tests/test_oracle_golden.py``).

Dropout is not restated: parity is run with dropout disabled.
"""
import math

import torch
import torch.nn.functional as F

ARCH = {  # utils/load_pretrained.py:103-109  -> depth, heads, dim, mlp_dim
    'vit-t16': (12, 3, 192, 768),
    's16': (12, 6, 384, 1536),
    'vit-s16': (12, 6, 384, 1536),
    'vit-b16': (12, 12, 768, 3072),
    'vit-l16': (24, 16, 1024, 4096),
}


def mapping_vit(backbone):
    """utils/load_pretrained.py:103-120."""
    if backbone is None:
        raise ValueError("Backbone must be specified.")
    if backbone.lower() not in ARCH:
        raise ValueError(f"Unsupported backbone: {backbone}.")
    return ARCH[backbone.lower()]


def layer_norm(x, w, b, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def quick_gelu(x):
    """model/gaviko.py:15-17."""
    return x * torch.sigmoid(1.702 * x)


def patchify(img, fp, ps):
    """Non-overlapping patches in Conv3d weight order (model/gaviko.py:383-385,532-533).

    img (B,C,D,H,W) -> (B, N, C*fp*ps*ps); token = d*nh*nw + h*nw + w, k = c*fp*ps*ps + kd*ps*ps + kh*ps + kw.
    """
    B, C, D, H, W = img.shape
    nd, nh, nw = D // fp, H // ps, W // ps
    x = img.reshape(B, C, nd, fp, nh, ps, nw, ps)
    x = x.permute(0, 2, 4, 6, 1, 3, 5, 7)          # B nd nh nw C fp ps ps
    return x.reshape(B, nd * nh * nw, C * fp * ps * ps)


def patch_embed(sd, img, fp, ps, prefix=''):
    w = sd[prefix + 'conv_proj.0.weight']
    b = sd[prefix + 'conv_proj.0.bias']
    return patchify(img, fp, ps) @ w.reshape(w.shape[0], -1).t() + b


def window_allow(DHW, local_k):
    """Boolean (N,N) mask of allowed keys, model/gaviko.py:212-227.

    allowed  <=>  for every axis  j >= i - k//2  and  j <= i + k - 1 - k//2  (asymmetric for even k).
    """
    D, H, W = DHW
    idx = torch.arange(D * H * W)
    coords = torch.stack([idx // (H * W), (idx // W) % H, idx % W], dim=1)  # (N,3)
    allow = torch.ones(D * H * W, D * H * W, dtype=torch.bool)
    for ax, k in enumerate(local_k):
        ci = coords[:, ax][:, None]
        cj = coords[:, ax][None, :]
        allow &= (cj >= ci - k // 2) & (cj <= ci + k - 1 - k // 2)
    return allow


def mhsa(x, w_norm, b_norm, w_qkv, w_out, b_out, heads, dim_head=64, qkv_hook=None, ln_hook=None):
    """model/vision_transformer.py:60-72 (no residual)."""
    B, T, _ = x.shape
    h = layer_norm(x, w_norm, b_norm)
    if ln_hook is not None:
        h = ln_hook(h)
    qkv = h @ w_qkv.t()
    if qkv_hook is not None:
        qkv = qkv_hook(qkv, h)
    q, k, v = qkv.chunk(3, dim=-1)
    q, k, v = (t.reshape(B, T, heads, dim_head).transpose(1, 2) for t in (q, k, v))
    a = torch.softmax(q @ k.transpose(-1, -2) * dim_head ** -0.5, dim=-1)
    o = (a @ v).transpose(1, 2).reshape(B, T, heads * dim_head)
    return o @ w_out.t() + b_out


def feed_forward(x, sd, p):
    """model/vision_transformer.py:26-38 (no residual); p = '...mlps.i.' or '...layers.i.1.'."""
    h = layer_norm(x, sd[p + 'net.0.weight'], sd[p + 'net.0.bias'])
    h = F.gelu(h @ sd[p + 'net.1.weight'].t() + sd[p + 'net.1.bias'])
    return h @ sd[p + 'net.4.weight'].t() + sd[p + 'net.4.bias']


def local_self_attention(loc, sd, p, dim, allow):
    """model/gaviko.py:229-244 (no residual); scale = dim**-0.5 (model/gaviko.py:201)."""
    z = layer_norm(loc, sd[p + 'norm.weight'], sd[p + 'norm.bias']) @ sd[p + 'proj_down.weight'].t() + sd[p + 'proj_down.bias']
    q, k, v = (z @ sd[p + 'qkv.weight'].t()).chunk(3, dim=-1)
    s = q @ k.transpose(-1, -2) * dim ** -0.5
    s = s.masked_fill(~allow.to(s.device), float('-inf'))
    a = torch.softmax(s, dim=-1)
    return (a @ v) @ sd[p + 'proj_up.weight'].t() + sd[p + 'proj_up.bias']


def awakening_prompt(g, loc, sd, p, P):
    """model/gaviko.py:149-187 (+ :20-47, :48-70, :84-119)."""
    wd, bd = sd[p + 'proj_down.0.weight'], sd[p + 'proj_down.0.bias']
    xl = quick_gelu(g @ wd.t() + bd)
    ll = quick_gelu(loc @ wd.t() + bd)
    r = xl.shape[-1]
    pl, cl, il = xl[:, :P], xl[:, P:P + 1], xl[:, P + 1:]
    ca = p + 'cls_analyzer.cls_analyzer_.'
    h = layer_norm(cl, sd[ca + '0.weight'], sd[ca + '0.bias'])
    h = F.gelu(h @ sd[ca + '1.weight'].t() + sd[ca + '1.bias'])
    imp = torch.sigmoid(h @ sd[ca + '3.weight'].t() + sd[ca + '3.bias'])          # (B,1,P)
    gb = p + 'gl_balancer.gl_balancer_.'
    gw = torch.sigmoid(layer_norm(cl, sd[gb + '0.weight'], sd[gb + '0.bias']) @ sd[gb + '1.weight'].t() + sd[gb + '1.bias'])  # (B,1,1)

    def xattn(tok, qp):
        q = pl @ sd[qp + 'weight'].t() + sd[qp + 'bias']
        a = torch.softmax(torch.einsum('bpd,bnd->bpn', q, tok) * r ** -0.5, dim=-1)
        return torch.einsum('bpn,bnd->bpd', a, tok)

    # NOTE the double slice: GlobalAttention.get_tokens slices an already sliced latent (model/gaviko.py:106-107,161,170)
    ctx_g = xattn(il[:, P + 1:], p + 'global_attention.query_proj.')
    ctx_l = xattn(ll, p + 'local_attention.query_proj.')
    fused = gw * ctx_g + (1 - gw) * ctx_l
    enh = fused * imp.transpose(1, 2)
    comb = torch.cat([enh, cl, il], dim=1)
    return comb @ sd[p + 'proj_up.weight'].t() + sd[p + 'proj_up.bias']


def gaviko_forward(sd, img, *, backbone, num_prompts, frame_patch_size, image_patch_size,
                   local_k, DHW, share_factor=1, dim_head=64, return_tokens=False):
    """model/gaviko.py:531-552 -> :291-306 -> :314-316, eval mode / dropout off."""
    depth, heads, dim, _ = mapping_vit(backbone)
    P = num_prompts
    e = patch_embed(sd, img, frame_patch_size, image_patch_size)
    B = e.shape[0]
    pos = sd['pos_embedding']
    g = torch.cat([
        (sd['prompt_embeddings'] + sd['prompt_positional_embedding']).expand(B, -1, -1),
        (sd['cls_token'] + pos[:, :1]).expand(B, -1, -1),
        e + pos[:, 1:]], dim=1)
    loc = e + pos[:, 1:]
    allow = window_allow(tuple(DHW), tuple(local_k))
    for i in range(depth):
        s = i // share_factor
        loc = local_self_attention(loc, sd, f'transformer.local_attns.{s}.', dim, allow) + loc
        a = f'transformer.attns.{i}.'
        g = mhsa(g, sd[a + 'norm.weight'], sd[a + 'norm.bias'], sd[a + 'to_qkv.weight'],
                 sd[a + 'to_out.0.weight'], sd[a + 'to_out.0.bias'], heads, dim_head) + g
        prompt = awakening_prompt(g, loc, sd, f'transformer.prompt_projs.{s}.', P)
        g = feed_forward(g, sd, f'transformer.mlps.{i}.') + g + prompt
    out = layer_norm(g, sd['transformer.norm.weight'], sd['transformer.norm.bias'])
    logits = out[:, :P + 1].mean(dim=1) @ sd['mlp_head.head.weight'].t() + sd['mlp_head.head.bias']
    if return_tokens:
        return logits, g, loc
    return logits


def focal_loss(logits, target, gamma=1.2, eps=1e-16, ignore_index=-100):
    """losses/focal_loss.py:84-111 — clamps the *logits*, softmaxes, clamps, softmaxes again; mean over non-ignored."""
    p1 = torch.softmax(torch.clamp(logits, eps, 1 - eps), dim=-1)
    p2 = torch.softmax(torch.clamp(p1, eps, 1 - eps), dim=-1)
    mask = target == ignore_index
    tgt = target * (~mask)
    pt = p2.gather(1, tgt.view(-1, 1)).squeeze(1) * (~mask)
    nll = (-torch.log(eps + pt)).masked_fill(mask, 0)
    loss = (1 - pt) ** gamma * nll
    return loss.sum() / (~mask).sum()


def cross_entropy(logits, target):
    """nn.CrossEntropyLoss alternative, train.py:179."""
    return F.cross_entropy(logits, target)


# ----------------------------------------------------------------------------------------------
# Variants.  Prefix handling: ViT blocks are 'transformer.layers.{i}.{0|1}.'.
# ----------------------------------------------------------------------------------------------

def _vit_tokens(sd, img, fp, ps, prefix=''):
    """model/vision_transformer.py:149-157: [cls ; patches] + pos."""
    e = patch_embed(sd, img, fp, ps, prefix)
    B = e.shape[0]
    x = torch.cat([sd[prefix + 'cls_token'].expand(B, -1, -1), e], dim=1)
    return x + sd[prefix + 'pos_embedding'][:, :x.shape[1]]


def _vit_block(x, sd, p, heads, dim_head, ff_index=1, qkv_hook=None):
    a = p + '0.'
    qw = sd.get(a + 'to_qkv.weight', sd.get(a + 'to_qkv.qkv.weight'))
    x = mhsa(x, sd[a + 'norm.weight'], sd[a + 'norm.bias'], qw, sd[a + 'to_out.0.weight'],
             sd[a + 'to_out.0.bias'], heads, dim_head, qkv_hook=qkv_hook) + x
    return x


def _pool_head(x, sd, pool, norm_p, head_p):
    x = layer_norm(x, sd[norm_p + 'weight'], sd[norm_p + 'bias'])
    x = x.mean(dim=1) if pool == 'mean' else x[:, 0]
    return x @ sd[head_p + 'weight'].t() + sd[head_p + 'bias']


def vit_forward(sd, img, *, backbone, frame_patch_size, image_patch_size, pool='cls', dim_head=64, prefix=''):
    """model/vision_transformer.py:149-164 (linear / bitfit / fft)."""
    depth, heads, dim, _ = mapping_vit(backbone)
    x = _vit_tokens(sd, img, frame_patch_size, image_patch_size, prefix)
    for i in range(depth):
        p = f'{prefix}transformer.layers.{i}.'
        x = _vit_block(x, sd, p, heads, dim_head)
        x = feed_forward(x, sd, p + '1.') + x
    return _pool_head(x, sd, pool, prefix + 'transformer.norm.', prefix + 'mlp_head.')


def adaptformer_forward(sd, img, *, backbone, frame_patch_size, image_patch_size, pool='cls', dim_head=64, scale=1.0):
    """model/adaptformer.py:58-78,93-99,194-209: x = ff(x) + x + up(relu(down(LN_a(x))))*scale."""
    depth, heads, dim, _ = mapping_vit(backbone)
    x = _vit_tokens(sd, img, frame_patch_size, image_patch_size)
    for i in range(depth):
        p = f'transformer.layers.{i}.'
        x = _vit_block(x, sd, p, heads, dim_head)
        ad = p + '1.'
        h = layer_norm(x, sd[ad + 'adapter_layer_norm_before.weight'], sd[ad + 'adapter_layer_norm_before.bias'])
        h = torch.relu(h @ sd[ad + 'down_adapter_proj.weight'].t() + sd[ad + 'down_adapter_proj.bias'])
        res = (h @ sd[ad + 'up_adapter_proj.weight'].t() + sd[ad + 'up_adapter_proj.bias']) * scale
        x = feed_forward(x, sd, p + '2.') + x + res
    return _pool_head(x, sd, pool, 'transformer.norm.', 'mlp_head.')


def melo_forward(sd, img, *, backbone, frame_patch_size, image_patch_size, r, alpha, pool='cls', dim_head=64):
    """model/melo.py:41-47: q += (alpha//r) B_q A_q h ; v += (alpha//r) B_v A_v h, h = LN output."""
    depth, heads, dim, _ = mapping_vit(backbone)
    pre = 'lora_vit.'
    x = _vit_tokens(sd, img, frame_patch_size, image_patch_size, pre)
    sc = alpha // r
    for i in range(depth):
        p = f'{pre}transformer.layers.{i}.'
        q = p + '0.to_qkv.'

        def hook(qkv, h, q=q):
            nq = (h @ sd[q + 'linear_a_q.weight'].t()) @ sd[q + 'linear_b_q.weight'].t()
            nv = (h @ sd[q + 'linear_a_v.weight'].t()) @ sd[q + 'linear_b_v.weight'].t()
            return torch.cat([qkv[..., :dim] + sc * nq, qkv[..., dim:-dim], qkv[..., -dim:] + sc * nv], dim=-1)

        x = _vit_block(x, sd, p, heads, dim_head, qkv_hook=hook)
        x = feed_forward(x, sd, p + '1.') + x
    return _pool_head(x, sd, pool, pre + 'transformer.norm.', pre + 'mlp_head.')


def ssf_forward(sd, img, *, backbone, frame_patch_size, image_patch_size, pool='cls', dim_head=64):
    """model/ssf.py:64-74,100-116,133-138,232-248."""
    depth, heads, dim, _ = mapping_vit(backbone)

    def ada(x, p, k):
        return x * sd[f'{p}ssf_scale_{k}'] + sd[f'{p}ssf_shift_{k}']

    e = patch_embed(sd, img, frame_patch_size, image_patch_size)
    e = ada(e, '', 1)                                                     # ssf.py:236 (channel-last after transpose)
    B = e.shape[0]
    x = torch.cat([sd['cls_token'].expand(B, -1, -1), e], dim=1)
    x = x + sd['pos_embedding'][:, :x.shape[1]]
    for i in range(depth):
        a = f'transformer.layers.{i}.0.'
        h = ada(layer_norm(x, sd[a + 'norm.weight'], sd[a + 'norm.bias']), a, 0)
        qkv = ada(h @ sd[a + 'to_qkv.weight'].t(), a, 1)
        T = x.shape[1]
        q, k, v = (t.reshape(B, T, heads, dim_head).transpose(1, 2) for t in qkv.chunk(3, dim=-1))
        att = torch.softmax(q @ k.transpose(-1, -2) * dim_head ** -0.5, dim=-1)
        o = (att @ v).transpose(1, 2).reshape(B, T, heads * dim_head)
        x = ada(o @ sd[a + 'to_out.0.weight'].t() + sd[a + 'to_out.0.bias'], a, 2) + x
        f = f'transformer.layers.{i}.1.'
        h = ada(layer_norm(x, sd[f + 'net.0.weight'], sd[f + 'net.0.bias']), f, 0)
        h = F.gelu(ada(h @ sd[f + 'net.1.weight'].t() + sd[f + 'net.1.bias'], f, 1))
        x = ada(h @ sd[f + 'net.4.weight'].t() + sd[f + 'net.4.bias'], f, 2) + x
    x = ada(layer_norm(x, sd['transformer.norm.weight'], sd['transformer.norm.bias']), 'transformer.', 1)
    x = x.mean(dim=1) if pool == 'mean' else x[:, 0]
    return x @ sd['mlp_head.weight'].t() + sd['mlp_head.bias']


def vpt_forward(sd, img, *, backbone, frame_patch_size, image_patch_size, deep_prompt, pool='cls', dim_head=64):
    """model/vpt.py:124-177.  Deep: layer i>=1 keeps x[:, 1+prompt_dim:] (prompt_dim, NOT num_prompts; vpt.py:151-153)."""
    depth, heads, dim, _ = mapping_vit(backbone)
    pre = 'vision_transformer.'
    x = _vit_tokens(sd, img, frame_patch_size, image_patch_size, pre)
    B = x.shape[0]
    pw, pb = sd['prompt_proj.weight'], sd['prompt_proj.bias']
    if not deep_prompt:
        pr = (sd['prompt_embeddings'] @ pw.t() + pb).expand(B, -1, -1)
        x = torch.cat([x[:, :1], pr, x[:, 1:]], dim=1)
    for i in range(depth):
        if deep_prompt:
            emb = sd['deep_prompt_embeddings'][i]
            pr = (emb @ pw.t() + pb).expand(B, -1, -1)
            cut = 1 if i == 0 else 1 + emb.shape[1]
            x = torch.cat([x[:, :1], pr, x[:, cut:]], dim=1)
        p = f'{pre}transformer.layers.{i}.'
        x = _vit_block(x, sd, p, heads, dim_head)
        x = feed_forward(x, sd, p + '1.') + x
    return _pool_head(x, sd, pool, pre + 'transformer.norm.', pre + 'mlp_head.')


def dvpt_share_mlp(x, sd, p, P, dim):
    """model/dvpt.py:36-47 (share_MLP.forward): QuickGELU BEFORE the down-projection, one prompt -> token cross-attention in the 20-wide
    latent with scale d_model ** -0.5 (dvpt.py:34), cls and token latents passed through, up-projection times the scalar gate."""
    z = quick_gelu(x) @ sd[p + 'prompt_key_proj_d.weight'].t() + sd[p + 'prompt_key_proj_d.bias']
    prompt, cls, tok = z[:, :P], z[:, P:P + 1], z[:, P + 1:]
    a = torch.softmax(prompt @ tok.transpose(-1, -2) * dim ** -0.5, dim=-1)
    out = torch.cat([a @ tok, cls, tok], dim=1)
    return (out @ sd[p + 'prompt_key_proj_u.weight'].t() + sd[p + 'prompt_key_proj_u.bias']) * sd[p + 'prompt_gate']


def dvpt_forward(sd, img, *, backbone, frame_patch_size, image_patch_size, num_prompts, pool='mean', dim_head=64):
    """model/dvpt.py:186-207 -> :49-63 (ResidualAttentionBlock) -> :65-82 (Transformer) — SURVEY.md §8 f3, oracle only so far.
    Tokens are [prompts ; cls ; patches] (dvpt.py:194-196); pool = 'mean' averages the normed rows 0..P (dvpt.py:81-82,201), pool = 'cls'
    takes row 0 of the normed sequence, which is the FIRST PROMPT row, not the cls row (reference behaviour)."""
    depth, heads, dim, _ = mapping_vit(backbone)
    P = num_prompts
    e = patch_embed(sd, img, frame_patch_size, image_patch_size)
    B = e.shape[0]
    x = torch.cat([sd['prompt_embeddings'].expand(B, -1, -1), sd['cls_token'].expand(B, -1, -1), e], dim=1)
    x = x + torch.cat([sd['prompt_positional_embedding'], sd['pos_embedding']], dim=1)
    for i in range(depth):
        p = f'transformer.layers.{i}.0.'
        a = p + 'attn.'
        x = mhsa(x, sd[a + 'norm.weight'], sd[a + 'norm.bias'], sd[a + 'to_qkv.weight'], sd[a + 'to_out.0.weight'], sd[a + 'to_out.0.bias'], heads, dim_head) + x
        prompt = dvpt_share_mlp(x, sd, p + 'prompt_proj.', P, dim)
        x = feed_forward(x, sd, p + 'mlp.') + x + prompt
    if pool == 'cls':
        y = layer_norm(x, sd['transformer.norm.weight'], sd['transformer.norm.bias'])[:, 0]
    else:
        y = layer_norm(x[:, :P + 1], sd['transformer.norm.weight'], sd['transformer.norm.bias']).mean(dim=1)
    return y @ sd['mlp_head.weight'].t() + sd['mlp_head.bias']


# ----------------------------------------------------------------------------------------------
# EVP: model/evp.py
# ----------------------------------------------------------------------------------------------

def evp_filter_plan(shape, rate):
    """What model/evp.py:124-146 (``PromptGenerator.fft``) does to a 5-D volume (B, C, D, H, W), written out:

    * ``fft2`` transforms the LAST TWO axes (H, W); ``fftshift`` / ``ifftshift`` without ``dim`` roll EVERY axis, batch included;
    * ``mask[:, :, w//2-line:w//2+line, h//2-line:h//2+line] = 1`` with ``w, h = x.shape[-2:]`` indexes axes 2 and 3 of the 5-D mask,
      i.e. DEPTH and HEIGHT (all of W), in shifted coordinates; ``line = int((w*h*rate) ** .5 // 2)``.

    In un-shifted coordinates the spectrum of slice d keeps every W frequency and loses the H frequencies kh with
    (kh + H//2) % H inside [H//2-line, H//2+line) — but only for the depth slices with (d + D//2) % D inside [W//2-line, W//2+line)
    (clipped to the axis); the other slices pass unchanged.  Returns (depth_hit [D] bool, freq_cut [H] bool)."""
    _, _, D, H, W = shape
    w, h = H, W                                   # evp.py:128 names them this way
    line = int((w * h * rate) ** .5 // 2)
    d_hit = torch.zeros(D, dtype=torch.bool)
    d_hit[w // 2 - line:w // 2 + line] = True      # the reference's own slice expressions (python slice semantics on each axis)
    k_cut = torch.zeros(H, dtype=torch.bool)
    k_cut[h // 2 - line:h // 2 + line] = True
    # un-shift: fftshift moves index i to (i + n//2) % n, so shifted position j holds un-shifted index (j - n//2) % n
    d_un = torch.zeros(D, dtype=torch.bool)
    d_un[(torch.arange(D) - D // 2) % D] = d_hit
    k_un = torch.zeros(H, dtype=torch.bool)
    k_un[(torch.arange(H) - H // 2) % H] = k_cut
    return d_un, k_un


def evp_filter_matrix(H, k_cut):
    """Real H x H matrix F with  real(ifft_H(fft_H(x) * (1 - cut))) = F @ x  for real x:  F = I - Re(IDFT diag(cut) DFT), in fp64."""
    n = torch.arange(H, dtype=torch.float64)
    k = n[k_cut.bool()]
    ang = 2 * math.pi * (n[:, None, None] - n[None, :, None]) * k[None, None, :] / H
    return torch.eye(H, dtype=torch.float64) - torch.cos(ang).sum(-1) / H


def evp_highpass(img, rate):
    """model/evp.py:124-146 restated with the same torch.fft calls (the closed form above is checked against this in the tests)."""
    mask = torch.zeros(img.shape)
    w, h = img.shape[-2:]
    line = int((w * h * rate) ** .5 // 2)
    mask[:, :, w // 2 - line:w // 2 + line, h // 2 - line:h // 2 + line] = 1
    f = torch.fft.fftshift(torch.fft.fft2(img, norm='forward')) * (1 - mask)
    return torch.fft.ifft2(torch.fft.ifftshift(f), norm='forward').real.abs()


def evp_highpass_closed_form(img, rate):
    """|F @ x| along H on the hit depth slices, |x| elsewhere (evp_filter_plan / evp_filter_matrix)."""
    d_un, k_un = evp_filter_plan(img.shape, rate)
    Fm = evp_filter_matrix(img.shape[3], k_un)
    y = img.double().clone()
    y[:, :, d_un] = torch.einsum('hk,bcdkw->bcdhw', Fm, y[:, :, d_un])
    return y.abs().float()


def evp_forward(sd, img, *, backbone, frame_patch_size, image_patch_size, pool='cls', dim_head=64, freq_nums=0.25):
    """model/evp.py:346-369 (ExplicitVisualPrompting.forward) -> :72-95 (PromptGenerator) -> :208-217 (Transformer).
    embedding_feature = Linear(conv_proj(img) tokens WITHOUT cls / pos); handcrafted_feature = a second Conv3d patch embedding of the
    high-passed volume; prompt_i = shared_mlp(GELU(lightweight_mlp_i(handcrafted + embedding))) is added to the patch rows (not cls)
    before every block; final LayerNorm, pool, plain Linear head."""
    depth, heads, dim, _ = mapping_vit(backbone)
    fp, ps = frame_patch_size, image_patch_size
    pg = 'prompt_generator.'
    wc, bc = sd['conv_proj.proj.weight'], sd['conv_proj.proj.bias']
    e = patchify(img, fp, ps) @ wc.reshape(wc.shape[0], -1).t() + bc                     # (B, N, dim)
    emb = e @ sd[pg + 'embedding_generator.weight'].t() + sd[pg + 'embedding_generator.bias']
    wh, bh = sd[pg + 'prompt_generator.proj.weight'], sd[pg + 'prompt_generator.proj.bias']
    hc = patchify(evp_highpass(img, freq_nums), fp, ps) @ wh.reshape(wh.shape[0], -1).t() + bh
    f = hc + emb
    B = e.shape[0]
    x = torch.cat([sd['cls_token'].expand(B, -1, -1), e], dim=1)
    x = x + sd['pos_embedding'][:, :x.shape[1]]
    for i in range(depth):
        lw = f'{pg}lightweight_mlp_{i}.0.'
        prompt = F.gelu(f @ sd[lw + 'weight'].t() + sd[lw + 'bias']) @ sd[pg + 'shared_mlp.weight'].t() + sd[pg + 'shared_mlp.bias']
        x = torch.cat([x[:, :1], prompt + x[:, 1:]], dim=1)
        p = f'transformer.layers.{i}.'
        x = _vit_block(x, sd, p, heads, dim_head)
        x = feed_forward(x, sd, p + '1.') + x
    return _pool_head(x, sd, pool, 'transformer.norm.', 'mlp_head.')
