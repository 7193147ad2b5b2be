"""GPU: the call sequence of the reference's scripts, restated line by line, on the real CUDA path.

The GPU box has no reference tree, so the unmodified scripts themselves run in tests/test_scripts_dropin.py (build container, kernels replaced
by their torch restatement); here the same model interaction runs against libgvk_sm100a.so: constructor kwargs arriving as OmegaConf-style
containers (train.py:112,518), `model.to(device)` then `model.to(device, dtype=torch.float32)` (train.py:155-157), the trainable-name list
(train.py:161-167), Adam over the trainable parameters + OneCycleLR (train.py:183-206), `model.train()`, forward / FocalLoss / backward /
`clip_grad_norm_(model.parameters(), 1.0)` / step (train.py:296-319), `.item()` reads (train.py:327-328), `model.eval()` + no_grad validation with
a `profile_macs`-style `torch.jit` trace of the model on the first batch (train.py:246-252,382-407), the trainable-only checkpoint
(train.py:478-483) and its reload through `load_vanilla_pretrain_with_adapters` + `load_state_dict(strict=False)` (eval.py:86-92), after which
the predictions are identical (eval.py:103-113).
"""
import contextlib
import io
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'script_shims'))
from omegaconf import DictConfig  # noqa: E402  (tests/script_shims: Sequence / Mapping containers that are not list / dict)

pytestmark = pytest.mark.gpu
DEVICE = 'cuda'

MODEL = dict(image_size=64, image_patch_size=16, frames=48, frame_patch_size=12, depth=12, heads=12, dim=768, mlp_dim=3072, dropout=0.1, emb_dropout=0.1,
             attn_drop=0.2, proj_drop=0.2, channels=1, num_classes=5, freeze_vit=True, pool='cls', backbone='vit-t16', num_prompts=8, prompt_latent_dim=20,
             local_dim=20, local_k=[3, 2, 2], DHW=[4, 4, 4], fp16=False, share_factor=1, r=4, alpha=4, prompt_dim=16, prompt_dropout=0.0)


def _build(method, tmp_path):
    from gaviko_b200.model.dvpt import DynamicVisualPromptTuning
    from gaviko_b200.model.evp import ExplicitVisualPrompting
    from gaviko_b200.model.gaviko import Gaviko
    from gaviko_b200.model.melo import MeLO
    from gaviko_b200.model.vision_transformer import VisionTransformer
    from gaviko_b200.model.vpt import PromptedVisionTransformer
    cfg = DictConfig(dict(model=dict(MODEL, method=method)))
    if method == 'deep_vpt':
        cfg['model']['deep_prompt'] = True
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            if method == 'gaviko':
                model = Gaviko(**cfg['model'])
            elif method == 'dvpt':
                model = DynamicVisualPromptTuning(**cfg['model'])
            elif method == 'evp':
                model = ExplicitVisualPrompting(**cfg['model'])
            elif method == 'melo':
                model = MeLO(vit=VisionTransformer(**cfg['model']), **cfg['model'])
            elif method == 'bitfit':
                model = VisionTransformer(**cfg['model'])
                for key, value in model.named_parameters():
                    value.requires_grad = ('bias' in key) or ('head' in key)
            else:
                model = PromptedVisionTransformer(**cfg['model'])
    finally:
        os.chdir(cwd)
    return model


@pytest.mark.parametrize('method', ['gaviko', 'dvpt', 'deep_vpt', 'melo', 'bitfit', 'evp'])
def test_train_eval_checkpoint_sequence(method, tmp_path):
    from gaviko_b200.losses.focal_loss import FocalLoss
    from gaviko_b200.utils.load_pretrained import load_vanilla_pretrain_with_adapters
    device = torch.device(DEVICE)
    torch.manual_seed(0)
    model = _build(method, tmp_path)
    model.to(device)
    model = model.to(device, dtype=torch.float32)
    tuning_params = [n for n, p in model.named_parameters() if p.requires_grad]
    criterion = FocalLoss(gamma=1.2)
    trainable = [p for p in model.parameters() if p.requires_grad]
    optimizer = torch.optim.Adam(trainable, lr=1e-4, eps=1e-8)
    steps = 4
    scheduler = torch.optim.lr_scheduler.OneCycleLR(optimizer, max_lr=3e-4, total_steps=steps, pct_start=0.3, div_factor=10.0, final_div_factor=1000.0,
                                                    anneal_strategy='cos', three_phase=False)
    g = torch.Generator().manual_seed(1)
    volumes = torch.rand(6, 1, 48, 64, 64, generator=g)
    labels = torch.tensor([0, 1, 2, 3, 4, 0])
    before = {n: p.detach().clone() for n, p in model.named_parameters() if p.requires_grad}
    traced_kinds = None
    for epoch in range(2):
        assert model.train() is None or method in ('melo', 'bitfit')
        for i in range(0, 4, 2):
            optimizer.zero_grad()
            inputs, y = volumes[i:i + 2].to(device, dtype=torch.float32), labels[i:i + 2].to(device)
            outputs = model(inputs)
            loss = criterion(outputs, y)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
            optimizer.step()
            scheduler.step()
            assert torch.isfinite(torch.tensor(loss.item()))
            _ = (torch.argmax(outputs, dim=1) == y).sum().item()
        model.eval()
        with torch.no_grad():
            inputs, y = volumes[4:6].to(device), labels[4:6].to(device)
            outputs = model(inputs)
            val_loss = criterion(outputs, y).item()
            if epoch == 0:       # calculate_flops -> torchprofile.profile_macs -> torch.jit._get_trace_graph(model, inputs)
                graph, _ = torch.jit._get_trace_graph(model, (inputs,), None)
                traced_kinds = [n.kind() for n in graph.nodes()]
                again = model(inputs)
                assert torch.equal(again, outputs), 'tracing must not disturb the model'
            assert val_loss == val_loss
    assert traced_kinds is not None and len(traced_kinds) > 0
    moved = [n for n, p in model.named_parameters() if p.requires_grad and not torch.equal(p.detach(), before[n])]
    assert len(moved) >= len(tuning_params) // 2, 'the optimiser should have moved the trainable tensors'
    # ---- trainable-only checkpoint and its reload over a fresh model with the same frozen backbone
    filtered = {k: v for k, v in model.state_dict().items() if k in tuning_params}
    assert sorted(filtered) == sorted(tuning_params)
    path = tmp_path / f'{method}_vit_t16_best_model_epoch2_acc0.5000.pt'
    torch.save(filtered, path)
    torch.manual_seed(0)
    fresh = _build(method, tmp_path)
    fresh.to(device)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        merged = load_vanilla_pretrain_with_adapters(MODEL['backbone'], dict(model=dict(MODEL)), str(path))
    finally:
        os.chdir(cwd)
    missing, unexpected = fresh.load_state_dict(merged, strict=False)
    assert not unexpected
    fresh.eval()
    model.eval()
    with torch.no_grad():
        x = volumes.to(device)
        a, b = model(x), fresh(x)
    assert torch.equal(a, b) and torch.equal(a.argmax(1), b.argmax(1))
