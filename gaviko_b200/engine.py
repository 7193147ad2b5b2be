"""Host-side orchestration of the GAViKO hot path over the C-ABI kernels (``gaviko_b200.ops``).

One ``torch.autograd.Function`` spans the whole model: forward launches the kernels layer by layer and keeps exactly the
activations the frozen-backbone backward needs; backward computes dX only through the frozen GEMMs / attention and dW only
for the trainable set (prompts, local attention, prompt fusion, head) — reference ``src/model/gaviko.py:291-306,531-552``
and the autograd graph ``src/train.py:305-311`` builds over it.  PyTorch is plumbing here (memory, streams, autograd glue).

Compute modes: 'fp32' (exact FFMA GEMMs + SIMT attention; the 1e-4 parity mode) and 'bf16' (tcgen05 GEMMs / attention with
fp32 accumulation; residual streams, LayerNorm statistics, softmax and every rank-r side path stay fp32).
"""
import os

import torch

from . import _lib as _L
from . import ops
from ._lib import GvkError

_SEED_MIX = 0x9E3779B97F4A7C15
_MASK63 = (1 << 63) - 1


def _resolve_dtype(compute_dtype, param_dtype):
    if compute_dtype is None:
        compute_dtype = os.environ.get('GAVIKO_COMPUTE_DTYPE') or None      # lets an unmodified script pick the mode (gaviko_b200/launch.py)
    if compute_dtype is None:
        return torch.float32 if param_dtype == torch.float32 else torch.bfloat16
    if isinstance(compute_dtype, torch.dtype):
        if compute_dtype in (torch.float32, torch.bfloat16):
            return compute_dtype
    elif str(compute_dtype).lower() in ('fp32', 'float32', 'f32'):
        return torch.float32
    elif str(compute_dtype).lower() in ('bf16', 'bfloat16'):
        return torch.bfloat16
    raise ValueError(f"compute_dtype must be 'fp32' or 'bf16', got {compute_dtype!r}")


_KEXT = 64   # K columns appended to the fc2 GEMM for the rank-r prompt up-projection (one 64-wide K block of the tcgen05 kernel)


class FrozenCache:
    """Compute-dtype copies (and transposes, for dgrad) of frozen tensors, rebuilt when the source tensor changes."""

    def __init__(self):
        self._store = {}

    def get(self, key, src, fn, extra=None):
        """`extra`: versions of further tensors the value depends on (the key stays stable, so a changed tensor REPLACES the entry)."""
        tag = (src.data_ptr(), src._version, src.device, src.dtype, extra)
        hit = self._store.get(key)
        if hit is not None and hit[0] == tag:
            return hit[1]
        with torch.no_grad():
            val = fn(src.detach())
        self._store[key] = (tag, val)
        return val

    def clear(self):
        self._store.clear()


def _f32(t):
    return t.detach().float().contiguous() if (t.dtype != torch.float32 or not t.is_contiguous()) else t.detach()


class GavikoEngine:
    def __init__(self, module, compute_dtype=None):
        self.__dict__['_module_ref'] = module          # plain attribute: the engine is not an nn.Module
        self._requested = compute_dtype
        self._cache = FrozenCache()
        self._step = 0
        # Optional {parameter name: fp32 tensor} of gradient accumulators (set by gaviko_b200.optim.FlatAdam): backward then adds
        # straight into the optimiser's flat exchange buffer and returns no per-tensor gradients to autograd.
        self.grad_sink = None

    # ------------------------------------------------------------------------------------------
    @property
    def module(self):
        return self._module_ref

    def set_compute_dtype(self, compute_dtype):
        _resolve_dtype(compute_dtype, torch.float32)
        self._requested = compute_dtype
        self._cache.clear()

    def compute_dtype(self):
        return _resolve_dtype(self._requested, self.module.pos_embedding.dtype)

    # ------------------------------------------------------------------------------------------
    def __call__(self, img):
        m = self.module
        _L.require_cuda(img)
        if m.pos_embedding.device != img.device:
            raise GvkError('model and input are on different devices')
        names, tensors = [], []
        for n, p in m.named_parameters():
            if p.requires_grad:
                names.append(n)
                tensors.append(p)
        need_grad = torch.is_grad_enabled() and len(tensors) > 0
        if need_grad:
            bad = [n for n in names if not ('head' in n or 'prompt' in n or 'local_attn' in n)]
            if bad:
                raise NotImplementedError('gaviko_b200 implements the frozen-backbone backward (freeze_vit=True); '
                                          f'backbone tensors require grad: {bad[:3]}...')
        out_dtype = img.dtype
        pause = _L.untraced()      # a jit trace (profile_macs in the reference's validation loop) cannot follow the kernels
        with pause, _L.device_guard(img):      # kernels launch on the CURRENT device's stream: make the input's device current (train.py never calls set_device)
            logits = _GavikoFn.apply(self, img, need_grad, names, *tensors)
        logits = pause.reattach(logits, img)
        return logits.to(out_dtype) if logits.dtype != out_dtype else logits

    # ------------------------------------------------------------------------------------------
    def _weights(self, cdt):
        """Frozen backbone tensors in compute dtype (+ transposes) and fp32 vectors; cached across calls."""
        m, c, cache = self.module, self.module._cfg, self._cache
        dim = c['dim']

        def mat(key, src):
            return cache.get((key, cdt), src, lambda t: t.reshape(t.shape[0], -1).to(cdt).contiguous())

        def mat_t(key, src):
            return cache.get((key, 't', cdt), src, lambda t: t.reshape(t.shape[0], -1).t().to(cdt).contiguous())

        def vec(key, src):
            return cache.get((key, 'v'), src, lambda t: t.float().contiguous())

        W = dict(conv_w=mat('conv_w', m.conv_proj[0].weight), conv_b=vec('conv_b', m.conv_proj[0].bias),
                 conv_w32=cache.get(('conv_w32',), m.conv_proj[0].weight, lambda t: t.reshape(t.shape[0], -1).float().contiguous()),
                 pos_patch=cache.get(('pos_patch',), m.pos_embedding, lambda t: t[0, 1:].float().contiguous()),
                 pos_cls=cache.get(('pos_cls',), m.pos_embedding, lambda t: t[0, :1].float().contiguous()),
                 cls=cache.get(('cls',), m.cls_token, lambda t: t.reshape(1, dim).float().contiguous()),
                 norm_w=vec('norm_w', m.transformer.norm.weight), norm_b=vec('norm_b', m.transformer.norm.bias), layers=[])
        for i in range(c['depth']):
            a, f = m.transformer.attns[i], m.transformer.mlps[i]
            W['layers'].append(dict(
                ln1_w=vec(('ln1w', i), a.norm.weight), ln1_b=vec(('ln1b', i), a.norm.bias),
                wqkv=mat(('wqkv', i), a.to_qkv.weight), wqkv_t=mat_t(('wqkv', i), a.to_qkv.weight),
                wo=mat(('wo', i), a.to_out[0].weight), wo_t=mat_t(('wo', i), a.to_out[0].weight), bo=vec(('bo', i), a.to_out[0].bias),
                ln2_w=vec(('ln2w', i), f.net[0].weight), ln2_b=vec(('ln2b', i), f.net[0].bias),
                w1=mat(('w1', i), f.net[1].weight), w1_t=mat_t(('w1', i), f.net[1].weight), b1=vec(('b1', i), f.net[1].bias),
                w2=mat(('w2', i), f.net[4].weight), w2_t=mat_t(('w2', i), f.net[4].weight), b2=vec(('b2', i), f.net[4].bias)))
            if cdt == torch.bfloat16:
                # fc2 weight with _KEXT extra K columns: the forward writes the prompt up-projection weight there every call, so that
                # fc2(act) + prompt comes out of ONE GEMM over [act | combined latent] (see forward)
                W['layers'][-1]['w2x'] = cache.get((('w2x', i), cdt), f.net[4].weight,
                                                   lambda t: torch.cat([t.to(cdt), torch.zeros(t.shape[0], _KEXT, device=t.device, dtype=cdt)], 1).contiguous())
        return W

    def _trainables(self):
        """fp32 views of the (possibly trainable) side-path tensors, keyed like the kernels expect."""
        m, c = self.module, self.module._cfg
        P, dim = c['num_prompts'], c['dim']
        T = dict(prompt_emb=_f32(m.prompt_embeddings).reshape(P, dim), prompt_pos=_f32(m.prompt_positional_embedding).reshape(P, dim),
                 head_w=_f32(m.mlp_head.head.weight), head_b=_f32(m.mlp_head.head.bias), local=[], fusion=[])
        for la in m.transformer.local_attns:
            T['local'].append(dict(ln_w=_f32(la.norm.weight), ln_b=_f32(la.norm.bias), wd=_f32(la.proj_down.weight), bd=_f32(la.proj_down.bias),
                                   wqkv=_f32(la.qkv.weight), wu=_f32(la.proj_up.weight), bu=_f32(la.proj_up.bias)))
        for pp in m.transformer.prompt_projs:
            est, bal = pp.cls_analyzer, pp.gl_balancer
            T['fusion'].append(dict(
                wd=_f32(pp.proj_down[0].weight), bd=_f32(pp.proj_down[0].bias), wu=_f32(pp.proj_up.weight), bu=_f32(pp.proj_up.bias),
                k=dict(wq_g=_f32(pp.global_attention.query_proj.weight), bq_g=_f32(pp.global_attention.query_proj.bias),
                       wq_l=_f32(pp.local_attention.query_proj.weight), bq_l=_f32(pp.local_attention.query_proj.bias),
                       a_ln_w=_f32(est[0].weight), a_ln_b=_f32(est[0].bias), a_w1=_f32(est[1].weight), a_b1=_f32(est[1].bias),
                       a_w3=_f32(est[3].weight), a_b3=_f32(est[3].bias),
                       g_ln_w=_f32(bal[0].weight), g_ln_b=_f32(bal[0].bias), g_w=_f32(bal[1].weight), g_b=_f32(bal[1].bias))))
        return T

    # parameter-name -> gradient-slot mapping (names are the checkpoint contract, reference train.py:161-167)
    def _grad_name_map(self):
        m = self.module
        out = {'prompt_embeddings': ('prompt_emb',), 'prompt_positional_embedding': ('prompt_pos',),
               'mlp_head.head.weight': ('head_w',), 'mlp_head.head.bias': ('head_b',)}
        for s in range(len(m.transformer.local_attns)):
            pre = f'transformer.local_attns.{s}.'
            for leaf, key in (('norm.weight', 'ln_w'), ('norm.bias', 'ln_b'), ('proj_down.weight', 'wd'), ('proj_down.bias', 'bd'),
                              ('qkv.weight', 'wqkv'), ('proj_up.weight', 'wu'), ('proj_up.bias', 'bu')):
                out[pre + leaf] = ('local', s, key)
        for s in range(len(m.transformer.prompt_projs)):
            pre = f'transformer.prompt_projs.{s}.'
            for leaf, key in (('proj_down.0.weight', 'wd'), ('proj_down.0.bias', 'bd'), ('proj_up.weight', 'wu'), ('proj_up.bias', 'bu')):
                out[pre + leaf] = ('fusion', s, key)
            for leaf, key in (('global_attention.query_proj.weight', 'wq_g'), ('global_attention.query_proj.bias', 'bq_g'),
                              ('local_attention.query_proj.weight', 'wq_l'), ('local_attention.query_proj.bias', 'bq_l'),
                              ('cls_analyzer.cls_analyzer_.0.weight', 'a_ln_w'), ('cls_analyzer.cls_analyzer_.0.bias', 'a_ln_b'),
                              ('cls_analyzer.cls_analyzer_.1.weight', 'a_w1'), ('cls_analyzer.cls_analyzer_.1.bias', 'a_b1'),
                              ('cls_analyzer.cls_analyzer_.3.weight', 'a_w3'), ('cls_analyzer.cls_analyzer_.3.bias', 'a_b3'),
                              ('gl_balancer.gl_balancer_.0.weight', 'g_ln_w'), ('gl_balancer.gl_balancer_.0.bias', 'g_ln_b'),
                              ('gl_balancer.gl_balancer_.1.weight', 'g_w'), ('gl_balancer.gl_balancer_.1.bias', 'g_b')):
                out[pre + leaf] = ('fusion', s, 'k', key)
        return out

    def _seed(self, layer, kind):
        base = (torch.initial_seed() * _SEED_MIX + self._step * 1315423911 + layer * 2654435761 + kind * 97) & _MASK63
        return base

    # ------------------------------------------------------------------------------------------
    def mhsa_fwd(self, qkv, B, T, H, D, dim):
        if qkv.dtype == torch.bfloat16 and D == 64:
            return ops.mhsa_fwd(qkv, B, T, H, D ** -0.5)          # tcgen05 flash attention
        return ops.attn_simt_fwd(qkv, B, T, H, D, q_off=0, k_off=dim, v_off=2 * dim, scale=D ** -0.5)

    def mhsa_bwd(self, qkv, o, lse, do, B, T, H, D, dim):
        if qkv.dtype == torch.bfloat16 and D == 64:
            return ops.mhsa_bwd(qkv, o, lse, do, B, T, H, D ** -0.5)
        return ops.attn_simt_bwd(qkv, o, lse, do, B, T, H, D, q_off=0, k_off=dim, v_off=2 * dim, scale=D ** -0.5)

    def forward(self, img, training_dropout, save):
        m, c = self.module, self.module._cfg
        cdt = self.compute_dtype()
        W, Tr = self._weights(cdt), self._trainables()
        B = img.shape[0]
        P, N, dim, H, D = c['num_prompts'], c['num_patches'], c['dim'], c['heads'], c['dim_head']
        T = P + 1 + N
        r_l, r_p = c['local_dim'], c['prompt_latent_dim']
        if img.dtype != torch.float32 or not img.is_contiguous():
            img = img.float().contiguous()
        if tuple(img.shape[1:]) != (c['channels'], c['grid'][0] * c['fp'], c['grid'][1] * c['ps'], c['grid'][2] * c['ps']):
            raise GvkError(f'unexpected volume shape {tuple(img.shape)}')

        # a1 + a2: patch gather -> GEMM whose epilogue adds bias + positional embedding and writes both token streams
        g = torch.empty((B * T, dim), device=img.device, dtype=torch.float32)
        loc = torch.empty((B * N, dim), device=img.device, dtype=torch.float32)
        fused = False
        if cdt != torch.float32:
            # bf16 mode: ONE kernel — TMA patch gather into a tf32 tcgen05 GEMM, bias + positional embedding + both token streams in the epilogue
            fused = ops.patch_embed(img, c['fp'], c['ps'], W['conv_w32'], W['conv_b'], W['pos_patch'], g, T, P + 1, out2=loc)
        if not fused:
            patches = ops.patch_gather(img, c['fp'], c['ps'], cdt)
            ops.gemm(patches, W['conv_w'], bias=W['conv_b'], pos=W['pos_patch'], rows_per_batch=N, out_batch_rows=T, out_row_offset=P + 1, out=g, out2=loc)
            del patches
        ops.fill_rows(Tr['prompt_emb'], Tr['prompt_pos'], g, T, 0, B)
        ops.fill_rows(W['cls'], W['pos_cls'], g, T, P, B)

        pr = ops.PREC_TF32 if cdt != torch.float32 else ops.PREC_FP32     # arithmetic of the rank-r side products
        drop_attn = c['attn_drop'] if training_dropout else 0.0
        drop_proj = c['proj_drop'] if training_dropout else 0.0
        self._step += 1
        layers = []
        for i in range(c['depth']):
            s = i // c['share_factor']
            Lw, La, Fu = W['layers'][i], Tr['local'][s], Tr['fusion'][s]
            st = {}
            # ---- local branch (model/gaviko.py:229-244, residual :301)
            d = ops.rowproj_down(loc, La['wd'], La['bd'], ln=(La['ln_w'], La['ln_b']), w2=La['wqkv'], prec=pr)
            seed_a, seed_p = self._seed(i, 1), self._seed(i, 2)
            ctx_l, lse_l = ops.attn_simt_fwd(d['z2'], B, N, 1, r_l, q_off=0, k_off=r_l, v_off=2 * r_l, scale=dim ** -0.5,
                                             window=c['local_k'], grid=c['DHW'], drop_p=drop_attn, seed=seed_a, prec=pr)
            fuse_ud = ops.rowproj_up_down_supported(dim, r_l, Fu['wd'].shape[0], pr)
            if fuse_ud:      # proj_up + proj_drop + residual and Awakening_Prompt.proj_down of the new local stream: one pass over the stream
                loc_new, dl = ops.rowproj_up_down(ctx_l, La['wu'], La['bu'], res=loc, up_drop_p=drop_proj, up_seed=seed_p, w2=Fu['wd'], bias2=Fu['bd'],
                                                  act=ops.ROWACT_QUICKGELU, save_pre=save)
            else:
                loc_new = ops.rowproj_up(ctx_l, La['wu'], La['bu'], res=loc, drop_p=drop_proj, seed=seed_p, prec=pr)
            # ---- frozen MHSA (model/vision_transformer.py:60-72, residual gaviko.py:302)
            h1, mean1, rstd1 = ops.layernorm_fwd(g, Lw['ln1_w'], Lw['ln1_b'], out_dtype=cdt, save_stats=save)
            qkv = ops.gemm(h1, Lw['wqkv'], out_dtype=cdt)
            del h1
            o, lse = self.mhsa_fwd(qkv, B, T, H, D, H * D)
            g_mid = ops.gemm(o, Lw['wo'], bias=Lw['bo'], res1=g)
            # ---- Awakening_Prompt (model/gaviko.py:149-187)
            fuse_ln2 = cdt == torch.bfloat16 and ops.layernorm_fwd_down_supported(g_mid, Fu['wd'].shape[0])
            if fuse_ln2:     # one pass over g_mid for its two readers: FeedForward's LayerNorm (below) and proj_down
                h2, mean2, rstd2, dg = ops.layernorm_fwd_down(g_mid, Lw['ln2_w'], Lw['ln2_b'], Fu['wd'], Fu['bd'], act=ops.ROWACT_QUICKGELU, save_pre=save,
                                                              save_stats=save)
            else:
                dg = ops.rowproj_down(g_mid, Fu['wd'], Fu['bd'], act=ops.ROWACT_QUICKGELU, save_pre=save, prec=pr)
            if not fuse_ud:
                dl = ops.rowproj_down(loc_new, Fu['wd'], Fu['bd'], act=ops.ROWACT_QUICKGELU, save_pre=save, prec=pr)
            comb, ll = dg['z'], dl['z']
            fsaved = ops.prompt_fusion_fwd(comb, ll, Fu['k'], B, T, N, P)      # comb: xl -> combined latent, in place
            # ---- frozen MLP (model/vision_transformer.py:26-38, residual + prompt gaviko.py:304)
            if not fuse_ln2:
                h2, mean2, rstd2 = ops.layernorm_fwd(g_mid, Lw['ln2_w'], Lw['ln2_b'], out_dtype=cdt, save_stats=save)
            hpre = torch.empty((B * T, c['mlp_dim']), device=img.device, dtype=cdt) if save else None
            mlp = c['mlp_dim']
            if 'w2x' in Lw and 3 * r_p <= _KEXT:
                # bf16 mode: g_out = g_mid + [act | comb] [W2 | Wu]^T + (b2 + bu) in one GEMM (K = mlp + 64).  The separate up-projection
                # pass (read g_mid, write g_mid + prompt: 406 MB per layer at B = 64) disappears.  The rank-r operands enter as bf16 hi / lo
                # pairs in three r-wide slots (hi*hi + lo*hi + hi*lo), so the trainable path keeps ~16 mantissa bits
                act_x = torch.empty((B * T, mlp + _KEXT), device=img.device, dtype=cdt)
                ops.gemm(h2, Lw['w1'], bias=Lw['b1'], act=ops.ACT_GELU_SAVE_GRAD if save else ops.ACT_GELU, aux=hpre, out=act_x[:, :mlp])   # aux = gelu'(pre)
                del h2
                ops.split_pack_bf16(comb, act_x[:, mlp:], 0b010)             # (hi, lo, hi)
                ops.split_pack_bf16(Fu['wu'], Lw['w2x'][:, mlp:], 0b100)     # (hi, hi, lo)
                g_out = ops.gemm(act_x, Lw['w2x'], bias=Lw['b2'] + Fu['bu'], res1=g_mid)
                del act_x
            else:
                g_tmp = ops.rowproj_up(comb, Fu['wu'], Fu['bu'], res=g_mid, prec=pr)
                act = ops.gemm(h2, Lw['w1'], bias=Lw['b1'], act=ops.ACT_GELU_SAVE_GRAD if save else ops.ACT_GELU, aux=hpre, out_dtype=cdt)   # aux = gelu'(pre)
                del h2
                g_out = ops.gemm(act, Lw['w2'], bias=Lw['b2'], res1=g_tmp)
                del act, g_tmp
            if save:
                st.update(loc_in=loc, mean_l=d['mean'], rstd_l=d['rstd'], z=d['z'], qkv_l=d['z2'], ctx_l=ctx_l, lse_l=lse_l, seed_a=seed_a, seed_p=seed_p,
                          g_in=g, mean1=mean1, rstd1=rstd1, qkv=qkv, o=o, lse=lse, g_mid=g_mid, mean2=mean2, rstd2=rstd2, hpre=hpre,
                          pre_g=dg['pre'], pre_l=dl['pre'], comb=comb, ll=ll, fsaved=fsaved, loc_out=loc_new)
                layers.append(st)
            g, loc = g_out, loc_new
        logits, pooled = ops.head_fwd(g, B, T, 0, P + 1, W['norm_w'], W['norm_b'], Tr['head_w'], Tr['head_b'])
        ctx = dict(layers=layers, g_final=g, pooled=pooled, B=B, T=T, N=N, W=W, Tr=Tr, cdt=cdt, drop_attn=drop_attn, drop_proj=drop_proj, names=[]) if save else None
        return logits, ctx

    # ------------------------------------------------------------------------------------------
    def backward(self, ctx, dlogits):
        c = self.module._cfg
        W, Tr, cdt = ctx['W'], ctx['Tr'], ctx['cdt']
        B, T, N = ctx['B'], ctx['T'], ctx['N']
        P, dim, H, D = c['num_prompts'], c['dim'], c['heads'], c['dim_head']
        r_l = c['local_dim']
        dev = dlogits.device
        lp = cdt != torch.float32
        pr = ops.PREC_TF32 if lp else ops.PREC_FP32
        # Gradient accumulators: views of the optimiser's flat buffer when a sink is attached, else of one zeroed flat buffer.
        nmap = self._grad_name_map()
        sink = self.grad_sink if (self.grad_sink is not None and all(n in self.grad_sink for n in ctx['names'])) else None
        G = dict(local=[dict() for _ in Tr['local']], fusion=[dict(k=dict()) for _ in Tr['fusion']])
        if sink is None:
            shapes = {n: _lookup(Tr, nmap[n]).shape for n in nmap}
            flat = torch.zeros(sum((s.numel() + 15) // 16 * 16 for s in shapes.values()), device=dev, dtype=torch.float32)
            off = 0
        for n, path in nmap.items():
            if sink is not None and n in sink:
                t = sink[n].view(_lookup(Tr, path).shape)
            elif sink is not None:
                t = torch.zeros_like(_lookup(Tr, path))        # frozen-by-user tensor: scratch accumulator
            else:
                k = shapes[n].numel()
                t = flat[off:off + k].view(shapes[n])
                off += (k + 15) // 16 * 16
            d = G
            for key in path[:-1]:
                d = d[key]
            d[path[-1]] = t
        ctx['used_sink'] = sink is not None

        # head (model/gaviko.py:306,314-316): only rows 0..P of each volume receive gradient
        dG = torch.zeros((B * T, dim), device=dev, dtype=torch.float32)
        dG_lp = torch.zeros((B * T, dim), device=dev, dtype=cdt) if lp else None
        ops.head_bwd(ctx['g_final'], B, T, 0, P + 1, W['norm_w'], W['norm_b'], Tr['head_w'], Tr['head_b'], ctx['pooled'], dlogits, dx=dG, dx_lp=dG_lp,
                     dwh=G['head_w'], dbh=G['head_b'])
        dLoc = None
        dcomb_next = None
        for i in reversed(range(c['depth'])):
            s = i // c['share_factor']
            Lw, La, Fu, st = W['layers'][i], Tr['local'][s], Tr['fusion'][s], ctx['layers'][i]
            gL, gF = G['local'][s], G['fusion'][s]
            a_in = dG_lp if lp else dG
            # ---- MLP dgrad: dH2 = (dG W2) * gelu'(hpre) W1
            dA = ops.gemm(a_in, Lw['w2_t'], act=ops.ACT_MUL_AUX, aux=st['hpre'], out_dtype=cdt)
            dH2 = ops.gemm(dA, Lw['w1_t'], out_dtype=cdt)     # bf16 mode: the gradient of the LayerNorm output travels as bf16 (half the bytes of
            del dA                                            # the GEMM's stores and of the LayerNorm backward's reads; its operands were bf16 anyway)
            # ---- prompt up-projection: d(comb) = dG Wu ; dWu, dbu
            if dcomb_next is None:
                dcomb = ops.rowproj_down(dG, Fu['wu'], transposed=True, prec=pr)['z']
            else:
                dcomb = dcomb_next       # projected from the output rows of the previous iteration's LayerNorm-backward pass
            ops.skinny_wgrad(st['comb'], dG, dw=gF['wu'], dw_layout='dr', dx_colsum=gF['bu'], prec=pr)
            dll = ops.prompt_fusion_bwd(st['comb'], st['ll'], dcomb, Fu['k'], st['fsaved'], gF['k'], B, T, N, P)
            du = ops.quickgelu_bwd(dcomb, st['pre_g'], out=dcomb)
            dul = ops.quickgelu_bwd(dll, st['pre_l'], out=dll)
            ops.skinny_wgrad(du, st['g_mid'], dw=gF['wd'], dw_layout='rd', da_colsum=gF['bd'], prec=pr)
            ops.skinny_wgrad(dul, st['loc_out'], dw=gF['wd'], dw_layout='rd', da_colsum=gF['bd'], prec=pr)
            # ---- d(g_mid) = dG + LN2'(dH2) + du Wd
            dGm_lp = torch.empty((B * T, dim), device=dev, dtype=cdt) if lp else None
            dGm = ops.layernorm_bwd(st['g_mid'], Lw['ln2_w'], st['mean2'], st['rstd2'], dy=dH2, dres=dG, dx=None if lp else dH2, dx_lp=dGm_lp, az=du, aw=Fu['wd'], prec=pr)
            del dH2
            # ---- d(loc_out) += dul Wd
            dctx = None
            if ops.rowproj_up_down_supported(dim, dul.shape[1], r_l, pr):
                # ... and the dgrad of LocalSelfAttention.proj_up (replayed proj_drop mask) from the rows just updated: one pass over d(loc)
                dLoc, dd = ops.rowproj_up_down(dul, Fu['wd'], transposed=True, res=dLoc, out=dLoc, w2=La['wu'], transposed2=True,
                                               dn_drop_p=ctx['drop_proj'], dn_seed=st['seed_p'])
                dctx = dd['z']
            elif dLoc is None:
                dLoc = ops.rowproj_up(dul, Fu['wd'], transposed=True, prec=pr)
            else:
                ops.rowproj_up(dul, Fu['wd'], transposed=True, res=dLoc, out=dLoc, prec=pr)
            # ---- MHSA dgrad
            dO = ops.gemm(dGm_lp if lp else dGm, Lw['wo_t'], out_dtype=cdt)
            dqkv = self.mhsa_bwd(st['qkv'], st['o'], st['lse'], dO, B, T, H, D, H * D)
            del dO
            dH1 = ops.gemm(dqkv, Lw['wqkv_t'], out_dtype=cdt)
            del dqkv
            dcomb_next = None
            if i > 0 and lp and ops.layernorm_bwd_down_supported(st['g_in'], Tr['fusion'][(i - 1) // c['share_factor']]['wu'].shape[1], pr):
                # the next iteration's d(comb) = dG Wu rides on this pass (its input rows are this pass's output rows)
                dG, dcomb_next = ops.layernorm_bwd(st['g_in'], Lw['ln1_w'], st['mean1'], st['rstd1'], dy=dH1, dres=dGm, dx=dGm, dx_lp=dG_lp, prec=pr,
                                                   ow=Tr['fusion'][(i - 1) // c['share_factor']]['wu'], ow_transposed=True)
            else:
                dG = ops.layernorm_bwd(st['g_in'], Lw['ln1_w'], st['mean1'], st['rstd1'], dy=dH1, dres=dGm, dx=dGm if lp else dH1, dx_lp=dG_lp)
            del dGm, dH1
            # ---- local branch backward
            if dctx is None:
                dctx = ops.rowproj_down(dLoc, La['wu'], transposed=True, drop_p=ctx['drop_proj'], seed=st['seed_p'], prec=pr)['z']
            ops.skinny_wgrad(st['ctx_l'], dLoc, dw=gL['wu'], dw_layout='dr', dx_colsum=gL['bu'], drop_p=ctx['drop_proj'], seed=st['seed_p'], prec=pr)
            dqkv_l = ops.attn_simt_bwd(st['qkv_l'], st['ctx_l'], st['lse_l'], dctx, B, N, 1, r_l, q_off=0, k_off=r_l, v_off=2 * r_l, scale=dim ** -0.5,
                                       window=c['local_k'], grid=c['DHW'], drop_p=ctx['drop_attn'], seed=st['seed_a'], prec=pr)
            ops.small_wgrad(dqkv_l, st['z'], gL['wqkv'])
            dz = ops.small_matmul(dqkv_l, La['wqkv'])
            ops.skinny_wgrad(dz, st['loc_in'], dw=gL['wd'], dw_layout='rd', da_colsum=gL['bd'], ln=(La['ln_w'], La['ln_b'], st['mean_l'], st['rstd_l']), prec=pr)
            dLoc = ops.layernorm_bwd(st['loc_in'], La['ln_w'], st['mean_l'], st['rstd_l'], dz=dz, w=La['wd'], dres=dLoc, dx=dLoc,
                                     dgamma=gL['ln_w'], dbeta=gL['ln_b'], prec=pr)
            ctx['layers'][i] = None     # release this layer's activations
        # prompt rows of the layer-0 input (model/gaviko.py:540-543): both prompt tensors receive the same gradient
        ops.batch_rowsum(dG, T, 0, P, B, out=G['prompt_emb'].view(P, dim), accumulate=True)
        ops.batch_rowsum(dG, T, 0, P, B, out=G['prompt_pos'].view(P, dim), accumulate=True)
        return G


def _lookup(G, path):
    v = G
    for k in path:
        v = v[k]
    return v


class _GavikoFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, img, need_grad, names, *tensors):
        m = engine.module
        # the only active dropouts under the reference's train() override are local_attns.*.attn_drop / proj_drop (gaviko.py:513-524)
        training_dropout = bool(m.transformer.local_attns[0].attn_drop.training) if len(m.transformer.local_attns) else False
        with torch.no_grad():
            logits, saved = engine.forward(img, training_dropout, need_grad)
        if saved is not None:
            saved['names'] = names
        ctx.engine, ctx.saved, ctx.names = engine, saved, names
        ctx.shapes = [(t.shape, t.dtype) for t in tensors]
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        if ctx.saved is None:
            raise RuntimeError('backward called on a forward that ran without gradient tracking')
        saved = ctx.saved
        with torch.no_grad(), _L.device_guard(dlogits):
            G = ctx.engine.backward(saved, dlogits.float().contiguous())
        ctx.saved = None
        if saved.get('used_sink'):
            return (None, None, None, None, *([None] * len(ctx.names)))     # already accumulated into the optimiser's flat buffer
        nmap = ctx.engine._grad_name_map()
        grads = []
        for n, (shape, dtype) in zip(ctx.names, ctx.shapes):
            grads.append(_lookup(G, nmap[n]).reshape(shape).to(dtype))
        return (None, None, None, None, *grads)
