"""the reference's UNMODIFIED scripts (src/train.py, src/eval.py, src/inference.py) run against the drop-in.

The scripts are executed as `__main__` by `gaviko_b200.launch` with `gaviko_b200/dropin` ahead of the reference's `src/` on `sys.path`, so
`from model.gaviko import Gaviko` & co resolve to gaviko_b200 while `utils.logging` / `data.dataset` stay the reference's own files; third-party
packages this image lacks come from tests/script_shims.  Exercised: OmegaConf containers as constructor kwargs (ListConfig local_k / DHW),
`model.to(device[, dtype])`, the trainable-name list, `model.train()` / `model.eval()`, forward + FocalLoss + backward + clip_grad_norm_ + Adam +
OneCycleLR, `profile_macs` jit-tracing the model inside the validation loop (train.py:246-252,405-407), the trainable-only checkpoint
(train.py:478-483), `load_vanilla_pretrain_with_adapters` + `load_state_dict(strict=False)` (eval.py:86-92, inference.py:86-92) and the result
files.  Without a GPU (this container) the kernels are replaced by their torch restatement (tests/ops_double.py); with one, the real CUDA path
runs.  The reference tree is located through GAVIKO_REFERENCE_SRC (default /root/reference/src) and the test is skipped where it is absent
(the GPU box): tests/test_script_equivalent_gpu.py restates the same call sequence there.
"""
import contextlib
import glob
import os
import sys

import numpy as np
import pandas as pd
import pytest
import torch

REF_SRC = os.environ.get('GAVIKO_REFERENCE_SRC', '/root/reference/src')
SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'script_shims')
pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(REF_SRC, 'train.py')), reason='reference scripts not present (set GAVIKO_REFERENCE_SRC)')

CONFIG = """{
  'utils':{ 'log_dir': '%(log)s', 'phase': 'train' },
  'data':{ 'batch_size': 2, 'num_workers': 0, 'data_path': '%(csv)s', 'image_folder': '%(img)s', 'test_data': '%(csv)s' },
  'model':{
      'image_size':64, 'image_patch_size':16, 'frames':48, 'frame_patch_size':12,
      'depth':12, 'heads':12, 'dim':768, 'mlp_dim':3072,        # unused by the constructors (the backbone name decides), as in gaviko.yaml:18-21
      'dropout':0.1, 'emb_dropout':0.1, 'attn_drop':0.2, 'proj_drop':0.2, 'channels':1, 'num_classes':5, 'freeze_vit':True, 'pool':'cls',
      'backbone': 'vit-t16', 'num_prompts':8, 'prompt_latent_dim':20, 'local_dim':20, 'local_k':[3,2,2], 'DHW':[4,4,4], 'fp16': False, 'share_factor':1,
      'r': 4, 'alpha': 4, 'prompt_dim': 16, 'prompt_dropout': 0.0,
  },
  'train':{
      'num_epochs': 2, 'lr': 1e-4, 'weight_decay': 1e-4, 'warmup_steps': 10, 'loss_fn': 'focal_loss', 'optimizer': 'adam', 'accumulation_steps': 1,
      'save_dir': '%(out)s', 'save_threshold': 0.0, 'fp16': False,      # train.fp16 is read by train.py:157 but missing from every shipped config
      'scheduler': { 'max_lr': 3e-4, 'pct_start': 0.3, 'div_factor': 10.0, 'final_div_factor': 1000.0, 'anneal_strategy': 'cos', 'three_phase': False },
      'patience': 15, 'deepspeed': { 'enabled': False, 'config': 'none' }, 'memory_verbose': False, 'flops_calculation': True,
  },
  'wandb':{ 'enable': False, 'project': 'gaviko', 'name': 'x', 'log_model': False, 'save_code': False, 'dir': '%(log)s' },
}"""


def _dataset(root):
    img = root / 'img'
    img.mkdir()
    rng = np.random.default_rng(0)
    rows = []
    for i in range(11):
        np.savez(img / f'vol{i}.npz', data=(rng.random((48, 64, 64), dtype=np.float32) * 900 + 50))      # raw intensities: RescaleIntensity maps them to [0, 1]
        rows.append(dict(mri_path=f'vol{i}.npz', kl_grade=i % 5, subset='train' if i < 4 else ('val' if i < 9 else 'test')))
    pd.DataFrame(rows).to_csv(root / 'data.csv', index=False)
    cfg = root / 'cfg.yaml'
    cfg.write_text(CONFIG % dict(log=str(root / 'log'), csv=str(root / 'data.csv'), img=str(img), out=str(root / 'weights')))
    return cfg


@contextlib.contextmanager
def _environment():
    import logging
    saved_path, saved_handlers = list(sys.path), list(logging.getLogger().handlers)
    sys.path.append(SHIMS)
    ctx = contextlib.nullcontext()
    if not torch.cuda.is_available():
        import ops_double
        ctx = ops_double.install()
    try:
        with ctx:
            yield
    finally:
        sys.path[:] = saved_path
        for h in list(logging.getLogger().handlers):
            if h not in saved_handlers:
                logging.getLogger().removeHandler(h)
                h.close()
        for name in [m for m in sys.modules if m.split('.')[0] in ('model', 'losses', 'utils', 'data', 'omegaconf', 'torchio', 'deepspeed', 'thop',
                                                                    'torchprofile', 'colorama')]:
            del sys.modules[name]


def _run(script, *argv):
    from gaviko_b200 import launch
    return launch.run(os.path.join(REF_SRC, script), list(argv))


@pytest.mark.parametrize('method', ['gaviko', 'dvpt', 'deep_vpt', 'melo', 'evp'])
def test_unmodified_train_eval_inference(method, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    cfg = _dataset(tmp_path)
    with _environment():
        ns = _run('train.py', '--config', str(cfg), '--method', method, '--results_dir', str(tmp_path / 'weights'))
        import torchprofile
        # ---- the scripts really imported the drop-in, not the reference's model files
        import model.gaviko as mg
        assert 'gaviko_b200' in os.path.realpath(mg.__file__) and mg.Gaviko.__module__.startswith('gaviko_b200.')
        import utils.logging as ul
        assert os.path.realpath(ul.__file__).startswith(os.path.realpath(REF_SRC))            # ... while the glue stays the reference's own
        assert len(torchprofile.CALLS) == 2, 'profile_macs should have traced the model once per epoch (train.py:405-407)'
        # ---- training log: one row per train / val step
        logs = glob.glob(str(tmp_path / 'log' / f'{method}_training_log_v1.csv'))
        assert len(logs) == 1
        log = pd.read_csv(logs[0])
        assert len(log) == 2 * (2 + 3) and np.isfinite(log['train_step_loss']).all() and (log['train_step_loss'] > 0).all()
        # ---- trainable-only checkpoint (train.py:478-483)
        ckpts = glob.glob(str(tmp_path / 'weights' / 'experiments' / method / f'{method}_vit_t16_best_model_epoch*_acc*.pt'))
        assert ckpts, 'no checkpoint written'
        ckpt = torch.load(sorted(ckpts)[-1], map_location='cpu')
        ref_names = _trainable_names(method, cfg)
        assert sorted(ckpt.keys()) == sorted(ref_names)
        # ---- eval.py and inference.py load it back (load_vanilla_pretrain_with_adapters -> load_state_dict(strict=False))
        _run('eval.py', '--config', str(cfg), '--method', method, '--checkpoint', sorted(ckpts)[-1], '--results_dir', str(tmp_path / 'eval'))
        res = pd.read_csv(tmp_path / 'eval' / f'{method}_vit_t16_eval_results_v1.csv')
        assert len(res) == 5 and set(res['outputs']) <= set(range(5))
        metrics = (tmp_path / 'eval' / f'{method}_vit_t16_eval_results_v1_metrics.txt').read_text()
        assert 'Test Accuracy' in metrics and 'Test AUC' in metrics
        _run('inference.py', '--config', str(cfg), '--method', method, '--checkpoint', sorted(ckpts)[-1], '--results_dir', str(tmp_path / 'infer'))
        res = pd.read_csv(tmp_path / 'infer' / f'{method}_vit_t16_inference_results_v1.csv')
        assert len(res) == 11 and set(res['outputs']) <= set(range(5))
    del ns


def _trainable_names(method, cfg):
    """Trainable parameter names of the REFERENCE's own class for this config (= what train.py:161-167 would collect with the reference model)."""
    import io
    from oracle import refload
    import yaml
    ref = refload.load()
    kw = yaml.safe_load(open(cfg))['model']
    kw['method'] = method
    if method in ('deep_vpt', 'shallow_vpt'):
        kw['deep_prompt'] = method == 'deep_vpt'
    with contextlib.redirect_stdout(io.StringIO()):
        if method == 'gaviko':
            m = ref.Gaviko(**kw)
        elif method == 'dvpt':
            m = ref.DynamicVisualPromptTuning(**kw)
        elif method == 'melo':
            m = ref.MeLO(vit=ref.VisionTransformer(**kw), **kw)
        elif method == 'evp':
            m = ref.ExplicitVisualPrompting(**kw)
        else:
            m = ref.PromptedVisionTransformer(**kw)
    for name in [k for k in sys.modules if k.split('.')[0] in ('model', 'losses', 'utils')]:
        del sys.modules[name]
    return [n for n, p in m.named_parameters() if p.requires_grad]
