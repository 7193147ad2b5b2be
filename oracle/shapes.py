"""Parameter names / shapes of the reference GAViKO model for a backbone, stated independently of both the reference code and
gaviko_b200 (so the CPU reference arm of bench.py touches no product code).  Mirrors reference model/gaviko.py:328-443.
TEST / BENCH INFRASTRUCTURE ONLY."""
import math

import torch

from .gaviko_oracle import mapping_vit


def gaviko_state_dict(backbone, *, num_prompts=32, num_patches=1000, patch_dim=3072, r_prompt=20, r_local=20, num_classes=5, share_factor=1, channels=1,
                      fp=12, ps=16):
    """Returns ({name: zeros tensor}, [trainable names]) in the reference's naming."""
    depth, heads, dim, mlp = mapping_vit(backbone)
    P = num_prompts
    sd = {'pos_embedding': (1, num_patches + 1, dim), 'cls_token': (1, 1, dim), 'prompt_positional_embedding': (1, P, dim), 'prompt_embeddings': (1, P, dim),
          'conv_proj.0.weight': (dim, channels, fp, ps, ps), 'conv_proj.0.bias': (dim,), 'transformer.norm.weight': (dim,), 'transformer.norm.bias': (dim,),
          'mlp_head.head.weight': (num_classes, dim), 'mlp_head.head.bias': (num_classes,)}
    for s in range(math.ceil(depth / share_factor)):
        p = f'transformer.local_attns.{s}.'
        sd.update({p + 'norm.weight': (dim,), p + 'norm.bias': (dim,), p + 'proj_down.weight': (r_local, dim), p + 'proj_down.bias': (r_local,),
                   p + 'qkv.weight': (3 * r_local, r_local), p + 'proj_up.weight': (dim, r_local), p + 'proj_up.bias': (dim,)})
        p = f'transformer.prompt_projs.{s}.'
        r = r_prompt
        sd.update({p + 'proj_down.0.weight': (r, dim), p + 'proj_down.0.bias': (r,), p + 'proj_up.weight': (dim, r), p + 'proj_up.bias': (dim,),
                   p + 'cls_analyzer.cls_analyzer_.0.weight': (r,), p + 'cls_analyzer.cls_analyzer_.0.bias': (r,),
                   p + 'cls_analyzer.cls_analyzer_.1.weight': (64, r), p + 'cls_analyzer.cls_analyzer_.1.bias': (64,),
                   p + 'cls_analyzer.cls_analyzer_.3.weight': (P, 64), p + 'cls_analyzer.cls_analyzer_.3.bias': (P,),
                   p + 'gl_balancer.gl_balancer_.0.weight': (r,), p + 'gl_balancer.gl_balancer_.0.bias': (r,),
                   p + 'gl_balancer.gl_balancer_.1.weight': (1, r), p + 'gl_balancer.gl_balancer_.1.bias': (1,),
                   p + 'global_attention.query_proj.weight': (r, r), p + 'global_attention.query_proj.bias': (r,),
                   p + 'local_attention.query_proj.weight': (r, r), p + 'local_attention.query_proj.bias': (r,)})
    for i in range(depth):
        a, f = f'transformer.attns.{i}.', f'transformer.mlps.{i}.'
        sd.update({a + 'norm.weight': (dim,), a + 'norm.bias': (dim,), a + 'to_qkv.weight': (3 * heads * 64, dim), a + 'to_out.0.weight': (dim, heads * 64),
                   a + 'to_out.0.bias': (dim,), f + 'net.0.weight': (dim,), f + 'net.0.bias': (dim,), f + 'net.1.weight': (mlp, dim), f + 'net.1.bias': (mlp,),
                   f + 'net.4.weight': (dim, mlp), f + 'net.4.bias': (dim,)})
    tensors = {k: torch.zeros(v) for k, v in sd.items()}
    trainable = [k for k in tensors if ('head' in k or 'prompt' in k or 'local_attn' in k)]      # freeze rule, model/gaviko.py:428-434
    return tensors, trainable
