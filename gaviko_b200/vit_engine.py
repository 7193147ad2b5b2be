"""Host-side orchestration of the ViT-family PEFT variants over the C-ABI kernels: ``--method linear | bitfit | adaptformer | melo |
ssf | shallow_vpt | deep_vpt`` of the reference (``src/train.py:111-153``).  All of them run the frozen blocks of
``src/model/vision_transformer.py:26-89`` with method-specific side paths:

* adaptformer — ``up(ReLU(down(LN_a(x))))`` in parallel to the MLP (``src/model/adaptformer.py:58-78,93-99``);
* melo        — rank-r LoRA on the q and v column blocks of ``to_qkv`` (``src/model/melo.py:41-47``);
* ssf         — ``x * scale + shift`` after the patch embedding, every LayerNorm and every Linear (``src/model/ssf.py:24-31,64-116,138,236``);
* vpt         — prompt tokens inserted after cls, once (shallow) or re-inserted at every layer with the reference's
                ``1 + prompt_dim`` slice (``src/model/vpt.py:124-161``), so the sequence length changes per layer;
* linear / bitfit — freeze rules only (``src/train.py:114-137``); bitfit needs the column sums of dY at every bias.
* evp         — ``x[:, 1:] += shared_mlp(GELU(lightweight_mlp_i(handcrafted + embedding)))`` before every block, where ``embedding`` is a
                Linear of the raw patch embedding and ``handcrafted`` a second patch embedding of the high-passed volume
                (``src/model/evp.py:70-95,124-146,208-217``); its rank dim / scale_factor needs general weight-gradient products (``ops.wgrad``).

Backward follows the freeze rule: dX through the frozen GEMMs / attention only as far as a trainable tensor needs it, dW only
for tensors with ``requires_grad``.  Same two compute modes as ``engine.GavikoEngine`` ('fp32' exact, 'bf16' tensor cores).
PyTorch is plumbing here (allocation, views, concatenation of token blocks, autograd glue); no arithmetic of the path runs in torch.
"""
import torch

from . import _lib as _L
from . import ops
from ._lib import GvkError
from .engine import FrozenCache, _f32, _resolve_dtype

_SEED_MIX = 0x9E3779B97F4A7C15
_MASK63 = (1 << 63) - 1


def _p(mod):
    """active dropout probability of an nn.Dropout (0 in eval mode)"""
    return float(mod.p) if (mod is not None and mod.training and mod.p > 0) else 0.0


class VitEngine:
    """kind: 'vit' (VisionTransformer: linear / bitfit / wrapped by MeLO), 'adaptformer', 'ssf', 'vpt', 'evp'."""

    def __init__(self, module, kind, compute_dtype=None):
        self.__dict__['_module_ref'] = module
        self.kind = kind
        self._requested = compute_dtype
        self._cache = FrozenCache()
        self._step = 0

    @property
    def module(self):
        return self._module_ref

    def set_compute_dtype(self, compute_dtype):
        _resolve_dtype(compute_dtype, torch.float32)
        self._requested = compute_dtype
        self._cache.clear()

    def compute_dtype(self):
        return _resolve_dtype(self._requested, self.vt.pos_embedding.dtype)

    @property
    def vt(self):
        m = self.module
        if self.kind == 'vpt':
            return m.vision_transformer
        if hasattr(m, 'lora_vit'):
            return m.lora_vit
        return m

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
            self._check_trainable(names)
        pause = _L.untraced()      # a jit trace (profile_macs in the reference's validation loop) cannot follow the kernels
        with pause, _L.device_guard(img):      # kernels launch on the CURRENT device's stream
            logits = _VitFn.apply(self, img, need_grad, names, *tensors)
        logits = pause.reattach(logits, img)
        return logits.to(img.dtype) if logits.dtype != img.dtype else logits

    def _check_trainable(self, names):
        """The engine computes gradients for the PEFT tensors only (biases, head, SSF, LoRA, adapter, prompts).  Any other tensor with
        requires_grad (--method fft, or AdaptFormer / SSF / VPT built with freeze_vit=False) would silently receive zeros: refuse instead."""
        key = tuple(names)
        if getattr(self, '_checked_names', None) == key:
            return
        W = self._weights(self.compute_dtype())
        ok = set()

        def add(v):
            if isinstance(v, (tuple, list)):
                for x in v:
                    add(x)
            elif isinstance(v, str):
                ok.add(v)
        add(list(W['names'].values()))
        for Lw in W['layers']:
            add(list(Lw['n'].values()))
        ok.update(('prompt_proj.weight', 'prompt_proj.bias', 'deep_prompt_embeddings', 'prompt_embeddings'))
        if self.kind == 'evp':
            ok.update(n for n in names if n.startswith('prompt_generator.'))
        bad = [n for n in names if n not in ok]
        if bad:
            raise NotImplementedError('gaviko_b200 implements the frozen-backbone backward of the PEFT methods (linear, bitfit, adaptformer, melo, ssf, '
                                      f'shallow_vpt, deep_vpt); these tensors require grad but have no weight-gradient kernel: {bad[:4]}'
                                      f'{" ..." if len(bad) > 4 else ""} ({len(bad)} tensors; full fine-tuning / freeze_vit=False is not supported)')
        self._checked_names = key

    def _seed(self, layer, kind):
        return (torch.initial_seed() * _SEED_MIX + self._step * 1315423911 + layer * 2654435761 + kind * 97) & _MASK63

    # ------------------------------------------------------------------------------------------
    def _layer_modules(self, i):
        layer = self.vt.transformer.layers[i]
        if self.kind == 'adaptformer':
            return layer[0], layer[2], layer[1]
        return layer[0], layer[1], None

    def _weights(self, cdt):
        """Per-layer tensors: GEMM operands in compute dtype (+ transposes for dgrad), vectors in fp32.  Cached on (data_ptr, version)."""
        vt, cache, c = self.vt, self._cache, self.module._cfg
        dim = c['dim']
        prefix = self._vt_prefix()

        def mat(key, src):
            return cache.get((key, cdt), src, lambda t: t.reshape(t.shape[0], -1).to(cdt).contiguous())

        def mat_t(key, src):
            return cache.get((key, 't', cdt), src, lambda t: t.reshape(t.shape[0], -1).t().to(cdt).contiguous())

        def vec(key, src):
            return None if src is None else cache.get((key, 'v'), src, lambda t: t.float().contiguous())

        conv = vt.conv_proj.proj if self.kind == 'evp' else vt.conv_proj[0]
        W = dict(conv_w=mat('conv_w', conv.weight), conv_b=vec('conv_b', conv.bias),
                 pos_patch=cache.get(('pos_patch',), vt.pos_embedding, lambda t: t[0, 1:].float().contiguous()),
                 pos_cls=cache.get(('pos_cls',), vt.pos_embedding, lambda t: t[0, :1].float().contiguous()),
                 cls=cache.get(('cls',), vt.cls_token, lambda t: t.reshape(1, dim).float().contiguous()),
                 norm_w=vec('norm_w', vt.transformer.norm.weight), norm_b=vec('norm_b', vt.transformer.norm.bias),
                 head_w=vec('head_w', vt.mlp_head.weight), head_b=vec('head_b', vt.mlp_head.bias), layers=[], names={})
        nm = W['names']
        nm['conv_b'] = prefix + ('conv_proj.proj.bias' if self.kind == 'evp' else 'conv_proj.0.bias')
        nm['norm_b'] = prefix + 'transformer.norm.bias'
        nm['head_w'], nm['head_b'] = prefix + 'mlp_head.weight', prefix + 'mlp_head.bias'
        ssf = self.kind == 'ssf'
        if ssf:
            W['ssf_patch'] = (vec('ssfp_s', vt.ssf_scale_1), vec('ssfp_h', vt.ssf_shift_1))
            W['ssf_final'] = (vec('ssff_s', vt.transformer.ssf_scale_1), vec('ssff_h', vt.transformer.ssf_shift_1))
            nm['ssf_patch'] = (prefix + 'ssf_scale_1', prefix + 'ssf_shift_1')
            nm['ssf_final'] = (prefix + 'transformer.ssf_scale_1', prefix + 'transformer.ssf_shift_1')
        for i in range(c['depth']):
            a, f, ad = self._layer_modules(i)
            ai, fi = (0, 2) if self.kind == 'adaptformer' else (0, 1)
            pa, pf = f'{prefix}transformer.layers.{i}.{ai}.', f'{prefix}transformer.layers.{i}.{fi}.'
            qkv_mod = a.to_qkv
            lora = hasattr(qkv_mod, 'linear_a_q')
            qkv_lin = qkv_mod.qkv if lora else qkv_mod
            Lw = dict(
                ln1_w=vec(('ln1w', i), a.norm.weight), ln1_b=vec(('ln1b', i), a.norm.bias),
                wqkv=mat(('wqkv', i), qkv_lin.weight), wqkv_t=mat_t(('wqkv', i), qkv_lin.weight),
                wo=mat(('wo', i), a.to_out[0].weight), wo_t=mat_t(('wo', i), a.to_out[0].weight), bo=vec(('bo', i), a.to_out[0].bias),
                ln2_w=vec(('ln2w', i), f.net[0].weight), ln2_b=vec(('ln2b', i), f.net[0].bias),
                w1=mat(('w1', i), f.net[1].weight), w1_t=mat_t(('w1', i), f.net[1].weight), b1=vec(('b1', i), f.net[1].bias),
                w2=mat(('w2', i), f.net[4].weight), w2_t=mat_t(('w2', i), f.net[4].weight), b2=vec(('b2', i), f.net[4].bias),
                drop_attn=a.dropout, drop_out=a.to_out[1], drop_ff1=f.net[3], drop_ff2=f.net[5],
                n=dict(ln1_b=pa + 'norm.bias', bo=pa + 'to_out.0.bias', ln2_b=pf + 'net.0.bias', b1=pf + 'net.1.bias', b2=pf + 'net.4.bias'))
            if ssf:
                Lw['ssf'] = dict(a0=(vec(('sa0s', i), a.ssf_scale_0), vec(('sa0h', i), a.ssf_shift_0)), a1=(vec(('sa1s', i), a.ssf_scale_1), vec(('sa1h', i), a.ssf_shift_1)),
                                 a2=(vec(('sa2s', i), a.ssf_scale_2), vec(('sa2h', i), a.ssf_shift_2)), f0=(vec(('sf0s', i), f.ssf_scale_0), vec(('sf0h', i), f.ssf_shift_0)),
                                 f1=(vec(('sf1s', i), f.ssf_scale_1), vec(('sf1h', i), f.ssf_shift_1)), f2=(vec(('sf2s', i), f.ssf_scale_2), vec(('sf2h', i), f.ssf_shift_2)))
                Lw['n'].update({k: (p_ + f'ssf_scale_{k[1]}', p_ + f'ssf_shift_{k[1]}') for k, p_ in (('a0', pa), ('a1', pa), ('a2', pa), ('f0', pf), ('f1', pf), ('f2', pf))})
                if cdt != torch.float32:
                    # bf16 mode: the SSF site behind a Linear folds into its operands, (x W^T + b) s + t = x (diag(s) W)^T + (b s + t), so the GEMM keeps
                    # a compile-time epilogue (bias [+ GELU]) instead of the generic one (4.5x slower on fc1).  The scales train: rebuilt when
                    # their version changes (torch's optimisers bump it; FlatAdam.step does so explicitly).  Backward is unchanged: it recovers
                    # the site's input from its saved output (ops.ssf_bwd) and runs the dgrad GEMM on the unscaled transposed weight.
                    def fold(key, wsrc, bsrc, site):
                        sc, sh = getattr(site[0], f'ssf_scale_{site[1]}'), getattr(site[0], f'ssf_shift_{site[1]}')
                        tag = (sc.data_ptr(), sc._version, sh.data_ptr(), sh._version, None if bsrc is None else (bsrc.data_ptr(), bsrc._version))
                        wf = cache.get((key, 'ssf_w', i, cdt), wsrc, lambda t: (t.float() * sc.detach().float()[:, None]).to(cdt).contiguous(), extra=tag)
                        bf = cache.get((key, 'ssf_b', i), wsrc, lambda t: ((0 if bsrc is None else bsrc.detach().float() * sc.detach().float()) + sh.detach().float()).contiguous(), extra=tag)
                        return wf, bf
                    Lw['fold'] = dict(a1=fold('wqkv', qkv_lin.weight, None, (a, 1)), a2=fold('wo', a.to_out[0].weight, a.to_out[0].bias, (a, 2)),
                                      f1=fold('w1', f.net[1].weight, f.net[1].bias, (f, 1)), f2=fold('w2', f.net[4].weight, f.net[4].bias, (f, 2)))
            if lora:
                r, s = qkv_mod.r, float(qkv_mod.alpha // qkv_mod.r)
                Aq, Av, Bq, Bv = qkv_mod.linear_a_q.weight, qkv_mod.linear_a_v.weight, qkv_mod.linear_b_q.weight, qkv_mod.linear_b_v.weight
                # stable keys + the versions of BOTH stacked tensors in the tag: an optimiser step replaces the entry instead of adding one
                tag = (Av.data_ptr(), Av._version)
                Lw['lora'] = dict(r=r, s=s,
                                  # LN1(x) @ (s A)^T gives the scaled latents; B and s B are the up weights of forward / the down weights of backward
                                  a_stack=cache.get(('loraA', i), Aq, lambda t: torch.cat([Aq.detach().float(), Av.detach().float()], 0).contiguous(), extra=tag),
                                  sa_stack=cache.get(('lorasA', i), Aq, lambda t: (s * torch.cat([Aq.detach().float(), Av.detach().float()], 0)).contiguous(), extra=tag),
                                  bq=_f32(Bq), bv=_f32(Bv),
                                  sbq=cache.get(('lorasBq', i), Bq, lambda t: (s * t.float()).contiguous()),
                                  sbv=cache.get(('lorasBv', i), Bv, lambda t: (s * t.float()).contiguous()))
                if cdt == torch.bfloat16 and 6 * r <= 128:
                    # bf16 mode: the LoRA update rides on the qkv GEMM as a K-extension (the scheme of GAViKO's prompt up-projection, engine.py): operand
                    # [LN1(x) | z] with z = s A LN1(x) in three 2r-wide hi / lo bf16 slots, weight [W_qkv | B] with B_q on the q rows, B_v on the v
                    # rows (hi*hi + lo*hi + hi*lo keeps ~16 mantissa bits).  The frozen part is cached; the forward rewrites the B columns every call
                    kext = 64 if 6 * r <= 64 else 128
                    Lw['lora']['kext'] = kext
                    Lw['lora']['wqkv_x'] = cache.get((('wqkvx', i, kext), cdt), qkv_lin.weight,
                                                     lambda t: torch.cat([t.to(cdt), torch.zeros(t.shape[0], kext, device=t.device, dtype=cdt)], 1).contiguous())
                pq = pa + 'to_qkv.'
                Lw['n'].update(aq=pq + 'linear_a_q.weight', av=pq + 'linear_a_v.weight', bq=pq + 'linear_b_q.weight', bv=pq + 'linear_b_v.weight')
            if ad is not None:
                if not isinstance(ad.scale, float) or ad.scale != 1.0 or ad.adapter_layernorm_option != 'in':
                    raise NotImplementedError('gaviko_b200 AdaptFormer supports the reference defaults (adapter_scalar="1.0", layernorm "in")')
                pd_ = f'{prefix}transformer.layers.{i}.1.'
                Lw['ad'] = dict(ln_w=_f32(ad.adapter_layer_norm_before.weight), ln_b=_f32(ad.adapter_layer_norm_before.bias),
                                wd=_f32(ad.down_adapter_proj.weight), bd=_f32(ad.down_adapter_proj.bias),
                                wu=_f32(ad.up_adapter_proj.weight), bu=_f32(ad.up_adapter_proj.bias), drop=float(ad.dropout))
                if ad.dropout != 0.0:
                    raise NotImplementedError('Adapter dropout > 0 is not implemented (reference default 0.0)')
                if cdt == torch.bfloat16:
                    # bf16 mode: the adapter's up-projection rides on the fc2 GEMM as a K-extension (hi / lo bf16 slots, as GAViKO's prompt path does):
                    # the frozen fc2 weight with kext zero columns is cached, the forward rewrites those columns from the up weight every call
                    kext = (3 * ad.down_dim + 63) // 64 * 64
                    Lw['ad']['kext'] = kext
                    Lw['ad']['w2x'] = cache.get((('ad_w2x', i, kext), cdt), f.net[4].weight,
                                                lambda t: torch.cat([t.to(cdt), torch.zeros(t.shape[0], kext, device=t.device, dtype=cdt)], 1).contiguous())
                Lw['n'].update(ad_ln_w=pd_ + 'adapter_layer_norm_before.weight', ad_ln_b=pd_ + 'adapter_layer_norm_before.bias',
                               ad_wd=pd_ + 'down_adapter_proj.weight', ad_bd=pd_ + 'down_adapter_proj.bias',
                               ad_wu=pd_ + 'up_adapter_proj.weight', ad_bu=pd_ + 'up_adapter_proj.bias')
            W['layers'].append(Lw)
        return W

    def _vt_prefix(self):
        if self.kind == 'vpt':
            return 'vision_transformer.'
        if hasattr(self.module, 'lora_vit'):
            return 'lora_vit.'
        return ''

    # ------------------------------------------------------------------------------------------
    def _attention(self, qkv, B, T, H, D, dim, drop_p, seed):
        if qkv.dtype == torch.bfloat16 and D == 64:      # tcgen05 flash attention, Philox dropout on the probabilities inside the softmax pass
            return ops.mhsa_fwd(qkv, B, T, H, D ** -0.5, drop_p=drop_p, seed=seed)
        return ops.attn_simt_fwd(qkv, B, T, H, D, q_off=0, k_off=dim, v_off=2 * dim, scale=D ** -0.5, drop_p=drop_p, seed=seed)

    def _attention_bwd(self, qkv, o, lse, do, B, T, H, D, dim, drop_p, seed):
        if qkv.dtype == torch.bfloat16 and D == 64:
            return ops.mhsa_bwd(qkv, o, lse, do, B, T, H, D ** -0.5, drop_p=drop_p, seed=seed)
        return ops.attn_simt_bwd(qkv, o, lse, do, B, T, H, D, q_off=0, k_off=dim, v_off=2 * dim, scale=D ** -0.5, drop_p=drop_p, seed=seed)

    def _prompts(self, i):
        """(P, dim) projected prompt tokens of layer i (model/vpt.py:127-131,146-153) and the embedding they came from."""
        m = self.module
        E = _f32(m.deep_prompt_embeddings[i] if m.deep_prompt else m.prompt_embeddings[0])
        wp, bp = _f32(m.prompt_proj.weight), _f32(m.prompt_proj.bias)
        if E.shape[1] % 4 == 0:
            # P rows only: the exact-fp32 GEMM (tiles over dim) instead of a rank-r row kernel, whose per-CTA panel staging (~0.1 ms) dwarfs the work
            return ops.gemm(E, wp, bias=bp), E
        return ops.rowproj_up(E, wp, bp), E

    def forward(self, img, save):
        m, vt, c = self.module, self.vt, self.module._cfg
        cdt = self.compute_dtype()
        lp = cdt != torch.float32
        pr = ops.PREC_TF32 if lp else ops.PREC_FP32
        W = self._weights(cdt)
        B = img.shape[0]
        N, dim, H, D = c['num_patches'], c['dim'], c['heads'], c['dim_head']
        if img.dtype != torch.float32 or not img.is_contiguous():
            img = img.float().contiguous()
        if tuple(img.shape[1:]) != (c['channels'], c['grid'][0] * c['fp'], c['grid'][1] * c['ps'], c['grid'][2] * c['ps']):
            raise GvkError(f'unexpected volume shape {tuple(img.shape)}')
        self._step += 1
        ssf = self.kind == 'ssf'
        # ---- patch embedding + [cls ; patches] + pos (model/vision_transformer.py:149-157), SSF site model/ssf.py:236
        T = N + 1
        patches = ops.patch_gather(img, c['fp'], c['ps'], cdt)
        x = torch.empty((B * T, dim), device=img.device, dtype=torch.float32)
        sp = W.get('ssf_patch', (None, None))
        ops.gemm(patches, W['conv_w'], bias=W['conv_b'], ssf_scale=sp[0], ssf_shift=sp[1], pos=W['pos_patch'], rows_per_batch=N, out_batch_rows=T,
                 out_row_offset=1, out=x)
        evp = self.kind == 'evp'
        if evp:
            ev = self._evp_setup(img, patches, W, cdt, B, N, T)
        del patches
        ops.fill_rows(W['cls'], W['pos_cls'], x, T, 0, B)
        ctx = dict(B=B, N=N, W=W, cdt=cdt, layers=[], x0=x if save else None, emb_drop=(_p(vt.dropout), self._seed(0, 9)))
        if evp and save:
            ctx['evp'] = ev
        if ctx['emb_drop'][0] > 0:
            x = ops.dropout(x, ctx['emb_drop'][0], ctx['emb_drop'][1])
        vpt = self.kind == 'vpt'
        p_prompt = _p(m.prompt_dropout) if vpt else 0.0      # model/vpt.py:57,129,148,152 (vpt.yaml ships prompt_dropout = 0.1)
        for i in range(c['depth']):
            Lw = W['layers'][i]
            st = dict(T_in=T)
            if evp:
                # prompt_i = shared_mlp(GELU(lightweight_mlp_i(f))) added to the patch rows (model/evp.py:85-95,210-214).  All B*T rows go through
                # the two GEMMs.  shared_mlp's bias rides on the GEMM as a K-extension: latent slot r of h holds 1 and column r of the padded weight
                # holds the bias, so zeroing the cls rows of h (B rows: plumbing) removes weight AND bias from them — and the same slot of the
                # weight gradient is the bias gradient.  The second GEMM then has the plain bias + fp32-residual epilogue
                E = ev['E']
                pre = torch.empty((B * T, E['rp']), device=img.device, dtype=cdt) if save else None
                h = ops.gemm(ev['f'], E['wi'][i], bias=E['bi'][i], act=ops.ACT_GELU, aux=pre, out_dtype=cdt)
                h[:, E['r']].fill_(1.0)
                h.view(B, T, -1)[:, 0].zero_()
                x = ops.gemm(h, E['ws'], bias=E['zero_dim'], res1=x)
                st.update(evp_h=h, evp_pre=pre)
            if vpt and (i == 0 or m.deep_prompt):
                # [cls ; P prompts ; rest]: at layers >= 1 the reference drops rows 1 .. prompt_dim (NOT 1 .. P), model/vpt.py:151-153
                pr_tok, E = self._prompts(i)
                P = pr_tok.shape[0]
                skip = 0 if i == 0 else m.deep_prompt_embeddings.shape[2]
                x3 = x.view(B, T, dim)
                pr_b = pr_tok.unsqueeze(0).expand(B, P, dim)
                seed_pr = self._seed(i, 7)
                if p_prompt > 0:     # the reference expands to the batch BEFORE the dropout: every volume gets its own mask
                    pr_b = ops.dropout(pr_b.reshape(B * P, dim).contiguous(), p_prompt, seed_pr).view(B, P, dim)
                x = torch.cat([x3[:, :1], pr_b, x3[:, 1 + skip:]], 1).reshape(-1, dim)
                st.update(vpt=(P, skip, E), vpt_drop=(p_prompt, seed_pr))
                T = x.shape[0] // B
            st['T'] = T
            sa = Lw.get('ssf', {})
            s_a0, s_a1, s_a2 = sa.get('a0', (None, None)), sa.get('a1', (None, None)), sa.get('a2', (None, None))
            s_f0, s_f1, s_f2 = sa.get('f0', (None, None)), sa.get('f1', (None, None)), sa.get('f2', (None, None))
            p_attn, p_out, p_ff1, p_ff2 = _p(Lw['drop_attn']), _p(Lw['drop_out']), _p(Lw['drop_ff1']), _p(Lw['drop_ff2'])
            seeds = [self._seed(i, k) for k in range(4)]
            # ---- attention (model/vision_transformer.py:60-72)
            lo = Lw.get('lora')
            if lo is not None and 'wqkv_x' in lo:
                # MeLO, bf16 mode: q / v updates B (s A LN1(x)) as a K-extension of the qkv GEMM (see _weights)
                r, kext = lo['r'], lo['kext']
                h1x = torch.empty((B * T, dim + kext), device=img.device, dtype=cdt)
                _, mean1, rstd1 = ops.layernorm_fwd(x, Lw['ln1_w'], Lw['ln1_b'], out=h1x[:, :dim], save_stats=save)
                z = ops.rowproj_down(x, lo['sa_stack'], ln=(Lw['ln1_w'], Lw['ln1_b']), prec=pr)['z']          # [M, 2r] = s * A LN1(x)
                bfull = torch.zeros((3 * dim, 2 * r), device=img.device, dtype=torch.float32)                # placement of the two up factors: plumbing
                bfull[:dim, :r] = lo['bq']
                bfull[2 * dim:, r:] = lo['bv']
                ops.split_pack_bf16(z, h1x[:, dim:], 0b010)                  # (hi, lo, hi)
                ops.split_pack_bf16(bfull, lo['wqkv_x'][:, dim:], 0b100)     # (hi, hi, lo)
                qkv = ops.gemm(h1x, lo['wqkv_x'], out_dtype=cdt)
                st['lora_z'] = z
                del h1x
            else:
                h1, mean1, rstd1 = ops.layernorm_fwd(x, Lw['ln1_w'], Lw['ln1_b'], out_dtype=cdt, ssf_scale=s_a0[0], ssf_shift=s_a0[1], save_stats=save)
                if lo is not None:
                    qkv32 = ops.gemm(h1, Lw['wqkv'])
                    z = ops.rowproj_down(x, lo['sa_stack'], ln=(Lw['ln1_w'], Lw['ln1_b']), prec=pr)['z']          # [M, 2r] = s * A LN1(x)
                    r = lo['r']
                    ops.rowproj_up(z[:, :r], lo['bq'], res=qkv32[:, :dim], out=qkv32[:, :dim], prec=pr)           # q += B_q (s A_q x)
                    ops.rowproj_up(z[:, r:], lo['bv'], res=qkv32[:, 2 * dim:], out=qkv32[:, 2 * dim:], prec=pr)  # v += B_v (s A_v x)
                    qkv = ops.cast_bf16(qkv32) if lp else qkv32
                    st['lora_z'] = z
                    del qkv32
                else:
                    fo = Lw.get('fold')
                    if fo is not None:
                        qkv = ops.gemm(h1, fo['a1'][0], bias=fo['a1'][1], out_dtype=cdt)
                    else:
                        qkv = ops.gemm(h1, Lw['wqkv'], ssf_scale=s_a1[0], ssf_shift=s_a1[1], out_dtype=cdt)
                del h1
            o, lse = self._attention(qkv, B, T, H, D, H * D, p_attn, seeds[0])
            if ssf or p_out > 0:
                fo = Lw.get('fold')
                y_a = ops.gemm(o, fo['a2'][0], bias=fo['a2'][1]) if fo is not None else ops.gemm(o, Lw['wo'], bias=Lw['bo'], ssf_scale=s_a2[0], ssf_shift=s_a2[1])
                x_mid = ops.dropout(y_a, p_out, seeds[1], res=x, out_dtype=torch.float32)
                st['y_a'] = y_a if ssf else None
            else:
                x_mid = ops.gemm(o, Lw['wo'], bias=Lw['bo'], res1=x)
            # ---- MLP (model/vision_transformer.py:26-38) + parallel adapter (model/adaptformer.py:93-99)
            h2, mean2, rstd2 = ops.layernorm_fwd(x_mid, Lw['ln2_w'], Lw['ln2_b'], out_dtype=cdt, ssf_scale=s_f0[0], ssf_shift=s_f0[1], save_stats=save)
            hpre = torch.empty((B * T, c['mlp_dim']), device=img.device, dtype=cdt) if save else None
            # Without an SSF site behind fc1 the backward needs only gelu'(pre): the forward epilogue saves the derivative (it shares the transcendental
            # with the activation) and the dgrad epilogue is one multiply instead of an erf.  SSF (model/ssf.py:77-80) needs the pre-activation itself.
            save_grad = save and 'f1' not in sa
            fo = Lw.get('fold')
            ad = Lw.get('ad')
            ad_fused = ad is not None and 'w2x' in ad and p_ff1 == 0 and p_ff2 == 0
            act_x = None
            if ad_fused:      # the activation goes straight into the left block of the K-extended fc2 operand (below)
                mlp, kext = c['mlp_dim'], ad['kext']
                act_x = torch.empty((B * T, mlp + kext), device=img.device, dtype=cdt)
            if fo is not None:
                act = ops.gemm(h2, fo['f1'][0], bias=fo['f1'][1], act=ops.ACT_GELU, aux=hpre, out_dtype=cdt)
            else:
                act = ops.gemm(h2, Lw['w1'], bias=Lw['b1'], ssf_scale=s_f1[0], ssf_shift=s_f1[1], act=ops.ACT_GELU_SAVE_GRAD if save_grad else ops.ACT_GELU,
                               aux=hpre, out_dtype=cdt, out=None if act_x is None else act_x[:, :mlp])
            st['gelu_grad_saved'] = save_grad
            del h2
            if p_ff1 > 0:
                act = ops.dropout(act, p_ff1, seeds[2])
            if ad_fused:
                # x_out = x_mid + [act | z] [W2 | Wu]^T + (b2 + bu): fc2 and the adapter's up-projection in ONE GEMM (K = mlp + kext)
                d = ops.rowproj_down(x_mid, ad['wd'], ad['bd'], ln=(ad['ln_w'], ad['ln_b']), act=ops.ROWACT_RELU, prec=pr)
                ops.split_pack_bf16(d['z'], act_x[:, mlp:], 0b010)           # (hi, lo, hi)
                ops.split_pack_bf16(ad['wu'], ad['w2x'][:, mlp:], 0b100)     # (hi, hi, lo)
                x_out = ops.gemm(act_x, ad['w2x'], bias=Lw['b2'] + ad['bu'], res1=x_mid)
                st.update(ad_z=d['z'], ad_mean=d['mean'], ad_rstd=d['rstd'])
                del act_x
            elif ssf or p_ff2 > 0:
                y_f = ops.gemm(act, fo['f2'][0], bias=fo['f2'][1]) if fo is not None else ops.gemm(act, Lw['w2'], bias=Lw['b2'], ssf_scale=s_f2[0], ssf_shift=s_f2[1])
                x_out = ops.dropout(y_f, p_ff2, seeds[3], res=x_mid, out_dtype=torch.float32)
                st['y_f'] = y_f if ssf else None
            else:
                x_out = ops.gemm(act, Lw['w2'], bias=Lw['b2'], res1=x_mid)
            del act
            if ad is not None and not ad_fused:
                d = ops.rowproj_down(x_mid, ad['wd'], ad['bd'], ln=(ad['ln_w'], ad['ln_b']), act=ops.ROWACT_RELU, prec=pr)
                self._up_chunked(d['z'], ad['wu'], ad['bu'], x_out, pr)
                st.update(ad_z=d['z'], ad_mean=d['mean'], ad_rstd=d['rstd'])
            if save:
                st.update(x_in=x, mean1=mean1, rstd1=rstd1, qkv=qkv, o=o, lse=lse, x_mid=x_mid, mean2=mean2, rstd2=rstd2, hpre=hpre,
                          drops=(p_attn, p_out, p_ff1, p_ff2), seeds=seeds)
                ctx['layers'].append(st)
            x = x_out
        pool = (0, 1) if vt.pool == 'cls' else (0, T)
        sf = W.get('ssf_final', (None, None))
        logits, pooled = ops.head_fwd(x, B, T, pool[0], pool[1], W['norm_w'], W['norm_b'], W['head_w'], W['head_b'], ssf_scale=sf[0], ssf_shift=sf[1])
        if not save:
            return logits, None
        ctx.update(x_final=x, pooled=pooled, T_final=T, pool=pool)
        return logits, ctx

    @staticmethod
    def _up_chunked(z, wu, bu, out, pr):
        """out += z @ wu^T + bu for a rank that may exceed what one staged [r, dim] panel allows (Adapter: r = 64)."""
        r, dim = z.shape[1], wu.shape[0]
        step = r if r * dim * 4 <= 160 * 1024 else 32
        for j0 in range(0, r, step):
            ops.rowproj_up(z[:, j0:j0 + step], wu[:, j0:j0 + step].contiguous(), bu if j0 == 0 else None, res=out, out=out, prec=pr)

    # ------------------------------------------------------------------------------------------
    def backward(self, ctx, dlogits, names):
        m, vt, c = self.module, self.vt, self.module._cfg
        W, cdt, B, N = ctx['W'], ctx['cdt'], ctx['B'], ctx['N']
        lp = cdt != torch.float32
        pr = ops.PREC_TF32 if lp else ops.PREC_FP32
        dim, H, D = c['dim'], c['heads'], c['dim_head']
        dev = dlogits.device
        params = dict(m.named_parameters())
        want = set(names)
        G = {n: torch.zeros(params[n].shape, device=dev, dtype=torch.float32) for n in names}

        def g(name):      # accumulator of a trainable tensor, or None when it is frozen
            return G[name] if name in want else None

        def g2(pair):
            return (g(pair[0]), g(pair[1])) if pair is not None else (None, None)

        nm = W['names']
        T = ctx['T_final']
        # which parts of the graph carry gradient to a trainable tensor?
        body_names = [n for n in names if 'mlp_head' not in n and n not in (nm['norm_b'],) and n not in nm.get('ssf_final', ())]
        need_body = len(body_names) > 0
        dX = torch.zeros((B * T, dim), device=dev, dtype=torch.float32) if need_body else None
        dX_lp = torch.zeros((B * T, dim), device=dev, dtype=cdt) if (need_body and lp) else None
        sf = W.get('ssf_final', (None, None))
        dsf = g2(nm.get('ssf_final'))
        ops.head_bwd(ctx['x_final'], B, T, ctx['pool'][0], ctx['pool'][1], W['norm_w'], W['norm_b'], W['head_w'], W['head_b'], ctx['pooled'], dlogits,
                     dx=dX, dx_lp=dX_lp, need_dx=need_body, ssf_scale=sf[0], ssf_shift=sf[1], dbeta=g(nm['norm_b']), dssf_scale=dsf[0], dssf_shift=dsf[1],
                     dwh=G.get(nm['head_w']) if nm['head_w'] in want else torch.zeros_like(W['head_w']),
                     dbh=G.get(nm['head_b']) if nm['head_b'] in want else torch.zeros_like(W['head_b']))
        if not need_body:
            return G
        vpt = self.kind == 'vpt'
        for i in reversed(range(c['depth'])):
            Lw, st = W['layers'][i], ctx['layers'][i]
            n_ = Lw['n']
            T = st['T']
            sa = Lw.get('ssf', {})
            p_attn, p_out, p_ff1, p_ff2 = st['drops']
            seeds = st['seeds']

            def site_bwd(dy, y_saved, key, bias_name, drop_p, seed, have_lp):
                """Gradient entering a [Linear -> SSF -> Dropout] site: replays the dropout mask, reduces the SSF / bias gradients and
                returns (d wrt the Linear output in compute dtype for the dgrad GEMM)."""
                d = dy
                fresh = False
                if drop_p > 0:
                    d = ops.dropout(d, drop_p, seed)
                    fresh = True
                if key in sa:
                    ds_, dh_ = g2(n_[key])
                    out = d if fresh else torch.empty_like(d)
                    ops.ssf_bwd(d, y=y_saved, scale=sa[key][0], shift=sa[key][1], dx=out, dscale=ds_, dshift=dh_)
                    d, fresh = out, True
                if g(bias_name) is not None:
                    ops.ssf_bwd(d, dshift=g(bias_name))
                if lp and d.dtype != cdt:
                    return have_lp if (have_lp is not None and not fresh) else ops.cast_bf16(d)
                return d

            # ---- MLP
            dY2 = site_bwd(dX, st.get('y_f'), 'f2', n_['b2'], p_ff2, seeds[3], dX_lp)
            dA = ops.gemm(dY2, Lw['w2_t'], act=ops.ACT_MUL_AUX if st['gelu_grad_saved'] else ops.ACT_GELU_BWD, aux=st['hpre'], out_dtype=cdt)
            if p_ff1 > 0:
                dA = ops.dropout(dA, p_ff1, seeds[2])
            if 'f1' in sa:
                ds_, dh_ = g2(n_['f1'])
                ops.ssf_bwd(dA, y=st['hpre'], scale=sa['f1'][0], shift=sa['f1'][1], dx=dA, dscale=ds_, dshift=dh_)
            if g(n_['b1']) is not None:
                ops.ssf_bwd(dA, dshift=g(n_['b1']))
            dH2 = ops.gemm(dA, Lw['w1_t'], out_dtype=cdt)     # bf16 mode: the gradient of a LayerNorm output travels as bf16 (half the bytes of the GEMM's
            del dA                                            # stores and of the LayerNorm backward's reads; its operands were bf16 anyway)
            f0 = sa.get('f0')
            dXm = ops.layernorm_bwd(st['x_mid'], Lw['ln2_w'], st['mean2'], st['rstd2'], dy=dH2, dres=dX, dx=None if lp else dH2, dbeta=g(n_['ln2_b']),
                                    beta=Lw['ln2_b'] if f0 else None, ssf_scale=f0[0] if f0 else None,
                                    dssf_scale=g2(n_['f0'])[0] if f0 else None, dssf_shift=g2(n_['f0'])[1] if f0 else None)
            ad = Lw.get('ad')
            if ad is not None:
                self._adapter_bwd(ad, n_, st, dX, dXm, g, pr, dX_lp if lp else None)
            # ---- attention
            dYa = site_bwd(dXm, st.get('y_a'), 'a2', n_['bo'], p_out, seeds[1], None)
            dO = ops.gemm(dYa, Lw['wo_t'], out_dtype=cdt)
            dqkv = self._attention_bwd(st['qkv'], st['o'], st['lse'], dO, B, T, H, D, H * D, p_attn, seeds[0])
            del dO
            if 'a1' in sa:
                ds_, dh_ = g2(n_['a1'])
                ops.ssf_bwd(dqkv, y=st['qkv'], scale=sa['a1'][0], shift=sa['a1'][1], dx=dqkv, dscale=ds_, dshift=dh_)
            lo = Lw.get('lora')
            dz = None
            if lo is not None:
                r = lo['r']
                # only the q and v column blocks feed the LoRA factors (model/melo.py:41-47): the rank-r kernels read fp32, so those two blocks are cast
                dq = ops.cast_f32(dqkv[:, :dim]) if lp else dqkv[:, :dim]
                dv = ops.cast_f32(dqkv[:, 2 * dim:]) if lp else dqkv[:, 2 * dim:]
                z = st['lora_z']
                if g(n_['bq']) is not None:
                    ops.skinny_wgrad(z[:, :r], dq, dw=g(n_['bq']), dw_layout='dr', prec=pr)
                if g(n_['bv']) is not None:
                    ops.skinny_wgrad(z[:, r:], dv, dw=g(n_['bv']), dw_layout='dr', prec=pr)
                dz = torch.cat([ops.rowproj_down(dq, lo['sbq'], transposed=True, prec=pr)['z'],
                                ops.rowproj_down(dv, lo['sbv'], transposed=True, prec=pr)['z']], 1)      # s * d(A LN1(x)), [M, 2r]
                ln = (Lw['ln1_w'], Lw['ln1_b'], st['mean1'], st['rstd1'])
                if g(n_['aq']) is not None:
                    ops.skinny_wgrad(dz[:, :r], st['x_in'], dw=g(n_['aq']), dw_layout='rd', ln=ln, prec=pr)
                if g(n_['av']) is not None:
                    ops.skinny_wgrad(dz[:, r:], st['x_in'], dw=g(n_['av']), dw_layout='rd', ln=ln, prec=pr)
                del dq, dv
            dH1 = ops.gemm(dqkv, Lw['wqkv_t'], out_dtype=cdt)
            del dqkv
            a0 = sa.get('a0')
            dX_lp = torch.empty((B * T, dim), device=dev, dtype=cdt) if lp else None
            dX = ops.layernorm_bwd(st['x_in'], Lw['ln1_w'], st['mean1'], st['rstd1'], dy=dH1, dz=dz, w=lo['a_stack'] if lo is not None else None,
                                   dres=dXm, dx=None if lp else dH1, dx_lp=dX_lp, dbeta=g(n_['ln1_b']), beta=Lw['ln1_b'] if a0 else None,
                                   ssf_scale=a0[0] if a0 else None, dssf_scale=g2(n_['a0'])[0] if a0 else None, dssf_shift=g2(n_['a0'])[1] if a0 else None)
            del dXm
            if 'vpt' in st:
                # undo the token insertion: prompt rows feed prompt_proj / the embeddings, dropped rows get zero gradient
                P, skip, E = st['vpt']
                T_in = st['T_in']
                p_prompt, seed_pr = st['vpt_drop']
                if p_prompt > 0:      # replay the forward's mask on the prompt rows before they are summed over the batch
                    dpr_b = ops.dropout(dX.view(B, T, dim)[:, 1:1 + P].reshape(B * P, dim).contiguous(), p_prompt, seed_pr)
                    dPr = ops.batch_rowsum(dpr_b, P, 0, P, B)
                else:
                    dPr = ops.batch_rowsum(dX, T, 1, P, B)                                               # [P, dim] = sum over the batch
                wp = _f32(m.prompt_proj.weight)
                pd = wp.shape[1]
                if g('prompt_proj.weight') is not None and pd % 4 == 0:
                    ops.wgrad(dPr, E, G['prompt_proj.weight'])                                           # [dim, prompt_dim] += dPr^T E (P rows, exact fp32)
                    if g('prompt_proj.bias') is not None:
                        ops.ssf_bwd(dPr, dshift=G['prompt_proj.bias'])
                elif g('prompt_proj.weight') is not None:
                    for j0 in range(0, pd, 32):
                        j1 = min(pd, j0 + 32)
                        ops.skinny_wgrad(E[:, j0:j1], dPr, dw=G['prompt_proj.weight'][:, j0:j1], dw_strides=(1, pd),
                                         dx_colsum=g('prompt_proj.bias') if j0 == 0 else None)
                elif g('prompt_proj.bias') is not None:
                    ops.ssf_bwd(dPr, dshift=G['prompt_proj.bias'])
                dE = ops.gemm(dPr, wp.t().contiguous()) if pd % 4 == 0 else ops.rowproj_down(dPr, wp, transposed=True)['z']   # [P, prompt_dim]
                ename = 'deep_prompt_embeddings' if m.deep_prompt else 'prompt_embeddings'
                if g(ename) is not None:
                    G[ename][i if m.deep_prompt else 0] += dE      # torch add of a [P, prompt_dim] tile: accumulation plumbing
                d3 = dX.view(B, T, dim)
                dX = torch.cat([d3[:, :1], torch.zeros((B, min(skip, T_in - 1), dim), device=dev, dtype=torch.float32), d3[:, 1 + P:]], 1).reshape(-1, dim)   # (a depleted sequence has fewer than `skip` rows to drop)
                assert dX.shape[0] == B * T_in
                dX_lp = ops.cast_bf16(dX) if lp else None
            if 'evp' in ctx:
                self._evp_layer_bwd(i, ctx, st, dX, dX_lp, pr)
            ctx['layers'][i] = None
            if not self._needs_below(i, names, nm):
                return G
        if 'evp' in ctx:
            self._evp_finish_bwd(ctx, G, want, pr)
        # ---- below layer 0: patch-embedding SSF site and the conv bias (bitfit)
        T0 = N + 1
        p_emb, seed_emb = ctx['emb_drop']
        if p_emb > 0:
            dX = ops.dropout(dX, p_emb, seed_emb)
        d_patch = dX[1:]     # logical row m -> physical row (m // N) * T0 + m % N of this view (cls rows skipped)
        if 'ssf_patch' in W:
            ds_, dh_ = g2(nm['ssf_patch'])
            ops.ssf_bwd(d_patch, y=ctx['x0'][1:], scale=W['ssf_patch'][0], shift=W['ssf_patch'][1], dscale=ds_, dshift=dh_, sub=W['pos_patch'],
                        rows_per_batch=N, batch_rows=T0, M=B * N)
        if g(nm['conv_b']) is not None:
            if 'ssf_patch' in W:
                raise NotImplementedError('conv bias gradient below an SSF site')
            ops.ssf_bwd(d_patch, dshift=G[nm['conv_b']], rows_per_batch=N, batch_rows=T0, M=B * N)
        return G

    def _needs_below(self, i, names, nm):
        """Does any trainable tensor live below layer i (so dX must keep flowing)?"""
        if i == 0:
            return True     # the post-loop block handles (and ignores) what is left
        prefix = self._vt_prefix()
        below = [f'{prefix}transformer.layers.{k}.' for k in range(i)]
        for n in names:
            if any(n.startswith(b) for b in below) or n in (nm['conv_b'],) or n in nm.get('ssf_patch', ()) or \
                    n.startswith('prompt_proj') or n.endswith('prompt_embeddings') or n.startswith('prompt_generator.'):
                return True
        return False

    # ------------------------------------------------------------------------------------------ EVP (model/evp.py)
    def _evp_filter(self, img):
        """PromptGenerator.fft (reference model/evp.py:124-146) in closed form, for a 5-D (B, C, D, H, W) volume.  ``fft2`` transforms the last
        two axes; ``fftshift`` / ``ifftshift`` without ``dim`` roll EVERY axis; ``mask[:, :, w//2-line:w//2+line, h//2-line:h//2+line] = 1`` with
        ``w, h = x.shape[-2:]`` hits axes 2 and 3 of the 5-D mask, i.e. DEPTH and HEIGHT, in shifted coordinates.  Un-shifted: on the depth
        slices d with (d + D//2) % D inside the first range, the H-frequencies k with (k + H//2) % H inside the second range are removed
        (every W-frequency is kept); other slices pass.  Removing a band along H, then ``.real``, is the real matrix
        F = I - (1/H) sum_{k cut} cos(2 pi k (n - m) / H) applied along H; ``abs`` follows.  Returns (F [H, H] fp32, hit [D] uint8) on the device."""
        import math
        _, _, D, H, Wd = img.shape
        rate = float(self.module.prompt_generator.freq_nums)
        key = ('evp_filter', D, H, Wd, rate, str(img.device))
        hit_f = getattr(self, '_evp_filter_cache', {}).get(key)
        if hit_f is None:
            w, h = H, Wd                                     # the names of evp.py:128
            line = int((w * h * rate) ** .5 // 2)
            d_s = torch.zeros(D, dtype=torch.bool)
            d_s[w // 2 - line:w // 2 + line] = True          # the reference's own slice expressions, python slice semantics per axis
            k_s = torch.zeros(H, dtype=torch.bool)
            k_s[h // 2 - line:h // 2 + line] = True
            d_un = torch.zeros(D, dtype=torch.bool)
            d_un[(torch.arange(D) - D // 2) % D] = d_s       # fftshift moves index i to (i + n // 2) % n
            k_un = torch.zeros(H, dtype=torch.bool)
            k_un[(torch.arange(H) - H // 2) % H] = k_s
            n = torch.arange(H, dtype=torch.float64)
            k = n[k_un]
            Fm = torch.eye(H, dtype=torch.float64) - torch.cos(2 * math.pi * (n[:, None, None] - n[None, :, None]) * k[None, None, :] / H).sum(-1) / H
            hit_f = (Fm.float().contiguous().to(img.device), d_un.to(torch.uint8).to(img.device))
            self.__dict__.setdefault('_evp_filter_cache', {})[key] = hit_f
        return hit_f

    def _evp_weights(self, cdt, T):
        """The prompt generator's tensors as GEMM operands: zero-padded to a latent width rp that is a multiple of 64 (the GEMM's K granularity)
        with at least one spare slot (slot r carries shared_mlp's bias, see forward), in compute dtype, plus the transposes the dgrad GEMMs
        read.  Trainable, so rebuilt every step (a few small casts: plumbing)."""
        m, c = self.module, self.module._cfg
        pg = m.prompt_generator
        r, dim, depth = c['evp_rank'], c['dim'], c['depth']
        rp = (r + 1 + 63) // 64 * 64
        dev = m.pos_embedding.device

        def padw(w, rows, cols):
            out = torch.zeros((rows, cols), device=dev, dtype=cdt)
            out[:w.shape[0], :w.shape[1]] = w.detach()
            return out

        def padv(v, n):
            out = torch.zeros(n, device=dev, dtype=torch.float32)
            out[:v.shape[0]] = v.detach()
            return out

        conv = pg.prompt_generator.proj
        lins = [getattr(pg, f'lightweight_mlp_{i}')[0] for i in range(depth)]
        ws = padw(pg.shared_mlp.weight, dim, rp)
        ws[:, r] = pg.shared_mlp.bias.detach()                 # bias as column r (h[:, r] = 1 on the patch rows, 0 on the cls rows)
        return dict(r=r, rp=rp, zero_dim=torch.zeros(dim, device=dev, dtype=torch.float32), we=padw(pg.embedding_generator.weight, rp, dim), be=padv(pg.embedding_generator.bias, rp),
                    wc=padw(conv.weight.reshape(r, -1), rp, conv.weight[0].numel()), bc=padv(conv.bias, rp),
                    ws=ws, ws_t=ws.t().contiguous(),
                    wi=[padw(l.weight, rp, rp) for l in lins], wi_t=[padw(l.weight.t(), rp, rp) for l in lins], bi=[padv(l.bias, rp) for l in lins])

    def _evp_setup(self, img, patches, W, cdt, B, N, T):
        """f = handcrafted + embedding (model/evp.py:70-79,346-351) as a [B*T, rp] matrix whose cls rows are zero."""
        c = self.module._cfg
        E = self._evp_weights(cdt, T)
        xraw = ops.gemm(patches, W['conv_w'], bias=W['conv_b'], out_dtype=cdt)              # conv_proj(img): no cls row, no positional embedding
        filt, hit = self._evp_filter(img)
        patches_hp = ops.patch_gather(ops.hfreq_filter(img, filt, hit), c['fp'], c['ps'], cdt)
        hc = ops.gemm(patches_hp, E['wc'], bias=E['bc'])                                   # second patch embedding (prompt_generator.proj), fp32
        f = torch.zeros((B * T, E['rp']), device=img.device, dtype=cdt)
        ops.gemm(xraw, E['we'], bias=E['be'], res1=hc, rows_per_batch=N, out_batch_rows=T, out_row_offset=1, out=f)
        return dict(E=E, xraw=xraw, patches_hp=patches_hp, f=f)

    def _evp_layer_bwd(self, i, ctx, st, dX, dX_lp, pr):
        """dX = gradient at the input of block i = gradient of prompt_i on the patch rows: shared_mlp / lightweight_mlp_i weight gradients and the
        running gradient of f.  Every product runs over all B*T rows: h and f have zero cls rows, bias sums skip the cls rows by row map."""
        ev = ctx['evp']
        E, B, N = ev['E'], ctx['B'], ctx['N']
        T = N + 1
        dev = dX.device
        if 'acc' not in ev:
            rp, dim = E['rp'], dX.shape[1]
            z = lambda *s: torch.zeros(s, device=dev, dtype=torch.float32)   # noqa: E731
            ev['acc'] = dict(ws=z(dim, rp), wi=[z(rp, rp) for _ in E['wi']], bi=[z(rp) for _ in E['wi']], we=z(rp, dim), bf=z(rp),
                             wc=z(rp, E['wc'].shape[1]))
        acc = ev['acc']
        dP = dX_lp if dX_lp is not None else dX
        ops.wgrad(dP, st['evp_h'], acc['ws'], prec=pr)            # column r (the ones slot of h) accumulates the bias gradient
        dpre = ops.gemm(dP, E['ws_t'], act=ops.ACT_GELU_BWD, aux=st['evp_pre'], out_dtype=ctx['cdt'])
        ops.wgrad(dpre, ev['f'], acc['wi'][i], prec=pr)
        ops.ssf_bwd(dpre[1:], dshift=acc['bi'][i], rows_per_batch=N, batch_rows=T, M=B * N)
        ev['df'] = ops.gemm(dpre, E['wi_t'][i], res1=ev.get('df'))      # fp32 [B*T, rp], summed over the layers

    def _evp_finish_bwd(self, ctx, G, want, pr):
        """d f feeds both feature generators (model/evp.py:70-79): embedding_generator (input: the raw patch embedding) and the handcrafted
        Conv3d (input: the patches of the high-passed volume).  Then the padded accumulators are cut to the parameter shapes."""
        ev = ctx['evp']
        E, acc, B, N = ev['E'], ev['acc'], ctx['B'], ctx['N']
        T, r = N + 1, ev['E']['r']
        df1 = ev['df'][1:]                                   # logical row m -> physical row (m // N) * T + m % N: the cls rows are skipped
        ops.ssf_bwd(df1, dshift=acc['bf'], rows_per_batch=N, batch_rows=T, M=B * N)
        ops.wgrad(df1, ev['xraw'], acc['we'], M=B * N, a_rows=(N, T), prec=pr)
        ops.wgrad(df1, ev['patches_hp'], acc['wc'], M=B * N, a_rows=(N, T), prec=pr)
        pg = 'prompt_generator.'
        out = {pg + 'shared_mlp.weight': acc['ws'][:, :r], pg + 'shared_mlp.bias': acc['ws'][:, r],
               pg + 'embedding_generator.weight': acc['we'][:r], pg + 'embedding_generator.bias': acc['bf'][:r],
               pg + 'prompt_generator.proj.weight': acc['wc'][:r], pg + 'prompt_generator.proj.bias': acc['bf'][:r]}
        for i in range(len(acc['wi'])):
            out[f'{pg}lightweight_mlp_{i}.0.weight'] = acc['wi'][i][:r, :r]
            out[f'{pg}lightweight_mlp_{i}.0.bias'] = acc['bi'][i][:r]
        for n, v in out.items():
            if n in want:
                G[n].copy_(v.reshape(G[n].shape))              # slice of a padded accumulator -> parameter shape: plumbing

    def _adapter_bwd(self, ad, n_, st, dX, dXm, g, pr, dX_lp=None):
        """x_out = ... + up(ReLU(down(LN_a(x_mid)))): gradients of the six adapter tensors and the contribution to d x_mid
        (model/adaptformer.py:58-78).  Rank 64 is processed in 32-wide slices (kernel limit of the rank-r gradient kernels)."""
        z = st['ad_z']
        r, dim = z.shape[1], dX.shape[1]
        ln = (ad['ln_w'], ad['ln_b'], st['ad_mean'], st['ad_rstd'])
        # bf16 mode, bottleneck a multiple of 64 (the reference's 64): the two [M, dim] x [dim, r] / [M, r] x [r, dim] products of the dgrad run on the
        # tcgen05 GEMM (one pass each over the gradient stream) instead of 32-wide slices through the rank-r row kernels (two passes each)
        tc = dX_lp is not None and r % 64 == 0
        cdt = dX_lp.dtype if tc else None
        dz = torch.empty_like(z)
        for j0 in range(0, r, 32):
            j1 = min(r, j0 + 32)
            if g(n_['ad_wu']) is not None:
                ops.skinny_wgrad(z[:, j0:j1], dX, dw=g(n_['ad_wu'])[:, j0:j1], dw_strides=(1, r), dx_colsum=g(n_['ad_bu']) if j0 == 0 else None, prec=pr)
            if not tc:
                dz[:, j0:j1] = ops.rowproj_down(dX, ad['wu'][:, j0:j1].contiguous(), transposed=True, prec=pr)['z']   # column-block copy: plumbing
        if tc:
            ops.gemm(dX_lp, ad['wu'].t().contiguous().to(cdt), out=dz)                  # dz = dX Wu  ([M, dim] x [dim, r])
        ops.relu_bwd(dz, z, out=dz)
        for j0 in range(0, r, 32):
            j1 = min(r, j0 + 32)
            dzc = dz[:, j0:j1]
            if g(n_['ad_wd']) is not None:
                ops.skinny_wgrad(dzc, st['x_mid'], dw=g(n_['ad_wd'])[j0:j1], dw_layout='rd', da_colsum=g(n_['ad_bd'])[j0:j1] if g(n_['ad_bd']) is not None else None,
                                 ln=ln, prec=pr)
            if not tc:
                ops.layernorm_bwd(st['x_mid'], ad['ln_w'], st['ad_mean'], st['ad_rstd'], dz=dzc, w=ad['wd'][j0:j1].contiguous(), dres=dXm, dx=dXm,
                                  dgamma=g(n_['ad_ln_w']), dbeta=g(n_['ad_ln_b']))
        if tc:
            dy = ops.gemm(ops.cast_bf16(dz), ad['wd'].t().contiguous().to(cdt), out_dtype=cdt)   # gradient of the adapter's LayerNorm output, [M, dim]
            ops.layernorm_bwd(st['x_mid'], ad['ln_w'], st['ad_mean'], st['ad_rstd'], dy=dy, dres=dXm, dx=dXm, dgamma=g(n_['ad_ln_w']), dbeta=g(n_['ad_ln_b']))


class _VitFn(torch.autograd.Function):
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
            G = ctx.engine.backward(saved, dlogits.float().contiguous(), ctx.names)
        return (None, None, None, None, *[G[n].reshape(s).to(d) for n, (s, d) in zip(ctx.names, ctx.shapes)])
