"""Host-side orchestration of the DVPT method (reference ``src/model/dvpt.py:36-63,186-207``) over the C-ABI kernels.

Per layer:  x = MHSA(x) + x;  prompt = share_MLP(x);  x = MLP(x) + x + prompt, with
``share_MLP(x) = (W_u cat[softmax(pl tok^T d_model^-0.5) tok ; cl ; tok] + b_u) * gate``, ``[pl ; cl ; tok] = W_d QuickGELU(x) + b_d``
(QuickGELU BEFORE the down-projection, one cross-attention, a scalar gate: the ungated ancestor of GAViKO's Awakening_Prompt).  The frozen
blocks run the same tcgen05 GEMM / attention kernels as ``engine.GavikoEngine``; the side path uses the rank-r row kernels plus the three
DVPT kernels of ``csrc/gvk_dvpt.cu``.  Backward follows the freeze rule of ``model/dvpt.py:157-163``: dX only through the frozen blocks, dW for
the prompts, every ``prompt_proj`` tensor and the head.  PyTorch is plumbing (allocation, autograd glue); no CPU fallback.
"""
import torch

from . import _lib as _L
from . import ops
from ._lib import GvkError
from .engine import FrozenCache, _f32, _resolve_dtype


class DvptEngine:
    def __init__(self, module, compute_dtype=None):
        self.__dict__['_module_ref'] = module
        self._requested = compute_dtype
        self._cache = FrozenCache()

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
        names, tensors = [], []
        for n, p in m.named_parameters():
            if p.requires_grad:
                names.append(n)
                tensors.append(p)
        need_grad = torch.is_grad_enabled() and len(tensors) > 0
        if need_grad:
            bad = [n for n in names if not ('prompt' in n or 'head' in n)]
            if bad:
                raise NotImplementedError('gaviko_b200 implements the frozen-backbone backward of DVPT (freeze_vit=True); '
                                          f'backbone tensors require grad: {bad[:3]}...')
        active = [d for d in [m.dropout] + [mod for mod in m.transformer.modules() if isinstance(mod, torch.nn.Dropout)] if d.training and d.p > 0]
        if active:
            raise NotImplementedError('DVPT with active backbone dropout is not implemented: call model.train() (with freeze_vit=True the reference puts the '
                                      'backbone in eval mode, model/dvpt.py:170-181) or model.eval() first')
        pause = _L.untraced()      # a jit trace (profile_macs in the reference's validation loop) cannot follow the kernels
        with pause, _L.device_guard(img):
            logits = _DvptFn.apply(self, img, need_grad, names, *tensors)
        logits = pause.reattach(logits, img)
        return logits.to(img.dtype) if logits.dtype != img.dtype else logits

    def _weights(self, cdt):
        m, c, cache = self.module, self.module._cfg, self._cache
        dim = c['dim']

        def mat(key, src):
            return cache.get((key, cdt), src, lambda t: t.reshape(t.shape[0], -1).to(cdt).contiguous())

        def mat_t(key, src):
            return cache.get((key, 't', cdt), src, lambda t: t.reshape(t.shape[0], -1).t().to(cdt).contiguous())

        def vec(key, src):
            return cache.get((key, 'v'), src, lambda t: t.float().contiguous())

        W = dict(conv_w=mat('conv_w', m.conv_proj[0].weight), conv_b=vec('conv_b', m.conv_proj[0].bias),
                 pos_patch=cache.get(('pos_patch',), m.pos_embedding, lambda t: t[0, 1:].float().contiguous()),
                 pos_cls=cache.get(('pos_cls',), m.pos_embedding, lambda t: t[0, :1].float().contiguous()),
                 cls=cache.get(('cls',), m.cls_token, lambda t: t.reshape(1, dim).float().contiguous()),
                 norm_w=vec('norm_w', m.transformer.norm.weight), norm_b=vec('norm_b', m.transformer.norm.bias), layers=[])
        for i in range(c['depth']):
            blk = m.transformer.layers[i][0]
            a, f = blk.attn, blk.mlp
            W['layers'].append(dict(
                ln1_w=vec(('ln1w', i), a.norm.weight), ln1_b=vec(('ln1b', i), a.norm.bias),
                wqkv=mat(('wqkv', i), a.to_qkv.weight), wqkv_t=mat_t(('wqkv', i), a.to_qkv.weight),
                wo=mat(('wo', i), a.to_out[0].weight), wo_t=mat_t(('wo', i), a.to_out[0].weight), bo=vec(('bo', i), a.to_out[0].bias),
                ln2_w=vec(('ln2w', i), f.net[0].weight), ln2_b=vec(('ln2b', i), f.net[0].bias),
                w1=mat(('w1', i), f.net[1].weight), w1_t=mat_t(('w1', i), f.net[1].weight), b1=vec(('b1', i), f.net[1].bias),
                w2=mat(('w2', i), f.net[4].weight), w2_t=mat_t(('w2', i), f.net[4].weight), b2=vec(('b2', i), f.net[4].bias)))
        return W

    def _side(self):
        """fp32 views of the (trainable) side-path tensors."""
        m, c = self.module, self.module._cfg
        P, dim = c['num_prompts'], c['dim']
        S = dict(prompt_emb=_f32(m.prompt_embeddings).reshape(P, dim), prompt_pos=_f32(m.prompt_positional_embedding).reshape(P, dim),
                 head_w=_f32(m.mlp_head.weight), head_b=_f32(m.mlp_head.bias), layers=[])
        for i in range(c['depth']):
            pp = m.transformer.layers[i][0].prompt_proj
            S['layers'].append(dict(wd=_f32(pp.prompt_key_proj_d.weight), bd=_f32(pp.prompt_key_proj_d.bias), wu=_f32(pp.prompt_key_proj_u.weight),
                                    bu=_f32(pp.prompt_key_proj_u.bias), gate=_f32(pp.prompt_gate)))
        return S

    def _names(self):
        out = {'prompt_embeddings': ('prompt_emb',), 'prompt_positional_embedding': ('prompt_pos',), 'mlp_head.weight': ('head_w',), 'mlp_head.bias': ('head_b',)}
        for i in range(self.module._cfg['depth']):
            pre = f'transformer.layers.{i}.0.prompt_proj.'
            for leaf, key in (('prompt_key_proj_d.weight', 'wd'), ('prompt_key_proj_d.bias', 'bd'), ('prompt_key_proj_u.weight', 'wu'),
                              ('prompt_key_proj_u.bias', 'bu'), ('prompt_gate', 'gate')):
                out[pre + leaf] = ('layers', i, key)
        return out

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _mhsa_fwd(qkv, B, T, H, D, dim):
        if qkv.dtype == torch.bfloat16 and D == 64:
            return ops.mhsa_fwd(qkv, B, T, H, D ** -0.5)
        return ops.attn_simt_fwd(qkv, B, T, H, D, q_off=0, k_off=dim, v_off=2 * dim, scale=D ** -0.5)

    @staticmethod
    def _mhsa_bwd(qkv, o, lse, do, B, T, H, D, dim):
        if qkv.dtype == torch.bfloat16 and D == 64:
            return ops.mhsa_bwd(qkv, o, lse, do, B, T, H, D ** -0.5)
        return ops.attn_simt_bwd(qkv, o, lse, do, B, T, H, D, q_off=0, k_off=dim, v_off=2 * dim, scale=D ** -0.5)

    def forward(self, img, save):
        m, c = self.module, self.module._cfg
        cdt = self.compute_dtype()
        W, Sd = self._weights(cdt), self._side()
        B = img.shape[0]
        P, N, dim, H, D, mlp = c['num_prompts'], c['num_patches'], c['dim'], c['heads'], c['dim_head'], c['mlp_dim']
        T = P + 1 + N
        if img.dtype != torch.float32 or not img.is_contiguous():
            img = img.float().contiguous()
        if tuple(img.shape[1:]) != (c['channels'], c['grid'][0] * c['fp'], c['grid'][1] * c['ps'], c['grid'][2] * c['ps']):
            raise GvkError(f'unexpected volume shape {tuple(img.shape)}')
        pr = ops.PREC_TF32 if cdt != torch.float32 else ops.PREC_FP32
        # tokens [prompts ; cls ; patches] + [prompt_pos ; pos] (model/dvpt.py:190-197)
        patches = ops.patch_gather(img, c['fp'], c['ps'], cdt)
        x = torch.empty((B * T, dim), device=img.device, dtype=torch.float32)
        ops.gemm(patches, W['conv_w'], bias=W['conv_b'], pos=W['pos_patch'], rows_per_batch=N, out_batch_rows=T, out_row_offset=P + 1, out=x)
        del patches
        ops.fill_rows(Sd['prompt_emb'], Sd['prompt_pos'], x, T, 0, B)
        ops.fill_rows(W['cls'], W['pos_cls'], x, T, P, B)
        scale = dim ** -0.5
        layers = []
        for i in range(c['depth']):
            Lw, Ls = W['layers'][i], Sd['layers'][i]
            h1, mean1, rstd1 = ops.layernorm_fwd(x, Lw['ln1_w'], Lw['ln1_b'], out_dtype=cdt, save_stats=save)
            qkv = ops.gemm(h1, Lw['wqkv'], out_dtype=cdt)
            del h1
            o, lse = self._mhsa_fwd(qkv, B, T, H, D, H * D)
            x_mid = ops.gemm(o, Lw['wo'], bias=Lw['bo'], res1=x)
            # ---- share_MLP (model/dvpt.py:36-47)
            a = ops.quickgelu_fwd(x_mid)
            comb = ops.rowproj_down(a, Ls['wd'], Ls['bd'], prec=pr)['z']
            del a
            pl, lse_x = ops.latent_xattn_fwd(comb, B, T, P, scale)          # comb: latents -> [attention output ; cls ; tokens], in place
            wg, bg = ops.gate_scale(Ls['wu'], Ls['gate']), ops.gate_scale(Ls['bu'], Ls['gate'])
            x_tmp = ops.rowproj_up(comb, wg, bg, res=x_mid, prec=pr)
            # ---- frozen MLP
            h2, mean2, rstd2 = ops.layernorm_fwd(x_mid, Lw['ln2_w'], Lw['ln2_b'], out_dtype=cdt, save_stats=save)
            hpre = torch.empty((B * T, mlp), device=img.device, dtype=cdt) if save else None
            act = ops.gemm(h2, Lw['w1'], bias=Lw['b1'], act=ops.ACT_GELU_SAVE_GRAD if save else ops.ACT_GELU, aux=hpre, out_dtype=cdt)
            del h2
            x_out = ops.gemm(act, Lw['w2'], bias=Lw['b2'], res1=x_tmp)
            del act, x_tmp
            if save:
                layers.append(dict(x_in=x, mean1=mean1, rstd1=rstd1, qkv=qkv, o=o, lse=lse, x_mid=x_mid, mean2=mean2, rstd2=rstd2, hpre=hpre,
                                   comb=comb, pl=pl, lse_x=lse_x, wg=wg))
            x = x_out
        # pool = 'mean': mean of the normed rows 0..P;  pool = 'cls': row 0 of the normed sequence = the FIRST PROMPT row (model/dvpt.py:78-82,201)
        pool = (0, P + 1) if m.pool == 'mean' else (0, 1)
        logits, pooled = ops.head_fwd(x, B, T, pool[0], pool[1], W['norm_w'], W['norm_b'], Sd['head_w'], Sd['head_b'])
        ctx = dict(layers=layers, x_final=x, pooled=pooled, pool=pool, B=B, T=T, W=W, Sd=Sd, cdt=cdt) if save else None
        return logits, ctx

    def backward(self, ctx, dlogits):
        c = self.module._cfg
        W, Sd, cdt, B, T = ctx['W'], ctx['Sd'], ctx['cdt'], ctx['B'], ctx['T']
        P, dim, H, D = c['num_prompts'], c['dim'], c['heads'], c['dim_head']
        dev = dlogits.device
        lp = cdt != torch.float32
        pr = ops.PREC_TF32 if lp else ops.PREC_FP32
        scale = dim ** -0.5
        z = lambda t: torch.zeros_like(t)  # noqa: E731
        G = dict(prompt_emb=z(Sd['prompt_emb']), prompt_pos=z(Sd['prompt_pos']), head_w=z(Sd['head_w']), head_b=z(Sd['head_b']),
                 layers=[{k: z(v) for k, v in Ls.items()} for Ls in Sd['layers']])
        dX = torch.zeros((B * T, dim), device=dev, dtype=torch.float32)
        dX_lp = torch.zeros((B * T, dim), device=dev, dtype=cdt) if lp else None
        ops.head_bwd(ctx['x_final'], B, T, ctx['pool'][0], ctx['pool'][1], W['norm_w'], W['norm_b'], Sd['head_w'], Sd['head_b'], ctx['pooled'], dlogits,
                     dx=dX, dx_lp=dX_lp, dwh=G['head_w'], dbh=G['head_b'])
        for i in reversed(range(c['depth'])):
            Lw, Ls, st, gL = W['layers'][i], Sd['layers'][i], ctx['layers'][i], G['layers'][i]
            # ---- MLP dgrad
            dA = ops.gemm(dX_lp if lp else dX, Lw['w2_t'], act=ops.ACT_MUL_AUX, aux=st['hpre'], out_dtype=cdt)
            dH2 = ops.gemm(dA, Lw['w1_t'])
            del dA
            # ---- share_MLP backward: prompt = comb (gate W_u)^T + gate b_u
            dcomb = ops.rowproj_down(dX, st['wg'], transposed=True, prec=pr)['z']
            dwg, dbg = torch.zeros_like(Ls['wu']), torch.zeros_like(Ls['bu'])
            ops.skinny_wgrad(st['comb'], dX, dw=dwg, dw_layout='dr', dx_colsum=dbg, prec=pr)
            ops.gate_grads(Ls['wu'], dwg, Ls['gate'], gL['wu'], gL['gate'])
            ops.gate_grads(Ls['bu'], dbg, Ls['gate'], gL['bu'], gL['gate'])
            dz = ops.latent_xattn_bwd(st['comb'], st['pl'], st['lse_x'], dcomb, B, T, P, scale)
            a = ops.quickgelu_fwd(st['x_mid'])                           # recomputed: saving it would cost [M, dim] fp32 per layer
            ops.skinny_wgrad(dz, a, dw=gL['wd'], dw_layout='rd', da_colsum=gL['bd'], prec=pr)
            del a
            da = ops.rowproj_up(dz, Ls['wd'], transposed=True, prec=pr)
            # ---- d x_mid = dX + LN2'(dH2) + da * QuickGELU'(x_mid)
            dXm = ops.layernorm_bwd(st['x_mid'], Lw['ln2_w'], st['mean2'], st['rstd2'], dy=dH2, dres=dX, dx=dH2)
            ops.quickgelu_bwd_add(da, st['x_mid'], res=dXm, out=dXm)
            del da
            # ---- MHSA dgrad
            dO = ops.gemm(ops.cast_bf16(dXm) if lp else dXm, Lw['wo_t'], out_dtype=cdt)
            dqkv = self._mhsa_bwd(st['qkv'], st['o'], st['lse'], dO, B, T, H, D, H * D)
            del dO
            dH1 = ops.gemm(dqkv, Lw['wqkv_t'])
            del dqkv
            dX_lp = torch.empty((B * T, dim), device=dev, dtype=cdt) if lp else None
            dX = ops.layernorm_bwd(st['x_in'], Lw['ln1_w'], st['mean1'], st['rstd1'], dy=dH1, dres=dXm, dx=dH1, dx_lp=dX_lp)
            del dXm
            ctx['layers'][i] = None
        # the prompt rows of the layer-0 input: both prompt tensors receive the same gradient (model/dvpt.py:194-196)
        ops.batch_rowsum(dX, T, 0, P, B, out=G['prompt_emb'], accumulate=True)
        ops.batch_rowsum(dX, T, 0, P, B, out=G['prompt_pos'], accumulate=True)
        return G


def _lookup(G, path):
    v = G
    for k in path:
        v = v[k]
    return v


class _DvptFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, img, need_grad, names, *tensors):
        with torch.no_grad():
            logits, saved = engine.forward(img, need_grad)
        ctx.engine, ctx.saved, ctx.names = engine, saved, names
        ctx.shapes = [(t.shape, t.dtype) for t in tensors]
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        if ctx.saved is None:
            raise RuntimeError('backward called on a forward that ran without gradient tracking')
        saved, ctx.saved = ctx.saved, None
        with torch.no_grad(), _L.device_guard(dlogits):
            G = ctx.engine.backward(saved, dlogits.float().contiguous())
        nmap = ctx.engine._names()
        return (None, None, None, None, *[_lookup(G, nmap[n]).reshape(s).to(d) for n, (s, d) in zip(ctx.names, ctx.shapes)])
