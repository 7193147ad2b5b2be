"""Named parity cases shared by the golden generator, the oracle tests and the GPU parity tests."""

FULL = dict(image_size=160, image_patch_size=16, frames=120, frame_patch_size=12)   # configs/gaviko.yaml:14-17
SMALL = dict(image_size=64, image_patch_size=16, frames=48, frame_patch_size=12)    # 4x4x4 = 64 tokens

GAVIKO_CASES = {
    # name: (ctor kwargs, batch)
    'gaviko_t16_full': (dict(FULL, num_classes=5, channels=1, freeze_vit=True, pool='cls', backbone='vit-t16',
                             num_prompts=32, prompt_latent_dim=20, local_dim=20, local_k=[6, 6, 6], DHW=[10, 10, 10],
                             dropout=0.0, emb_dropout=0.0, attn_drop=0.0, proj_drop=0.0, share_factor=1, fp16=False), 2),
    'gaviko_t16_small': (dict(SMALL, num_classes=5, channels=1, freeze_vit=True, pool='cls', backbone='vit-t16',
                              num_prompts=8, prompt_latent_dim=20, local_dim=20, local_k=[3, 2, 2], DHW=[4, 4, 4],
                              dropout=0.0, emb_dropout=0.0, attn_drop=0.0, proj_drop=0.0, share_factor=2, fp16=False), 3),
}

# Cases on the reference's OWN random initialisation (the north star's "same random-init ViT weights"): the reference constructor and the drop-in
# constructor consume the RNG identically (tests/test_dropin_surface.py::test_seeded_construction_equals_reference_init), so
# `torch.manual_seed(seed); Gaviko(**kw)` gives both sides the same weights without shipping them.  The golden file carries a fingerprint of
# every tensor (sum, sum of squares) so a drifted init is detected rather than reported as a parity failure.
# name: (ctor kwargs, batch, seed, store)   store 'all' = every trainable gradient, 'subset' = full gradients of the prompts / head / three
# layers + 32-element chunk sums of every other tensor (ViT-B / ViT-L gradients are 7 / 18 MB per case otherwise)
_G = dict(num_classes=5, channels=1, freeze_vit=True, pool='cls', prompt_latent_dim=20, local_dim=20, dropout=0.0, emb_dropout=0.0, attn_drop=0.0,
          proj_drop=0.0, fp16=False)
GAVIKO_INIT_CASES = {
    'gaviko_t16_small_init': (dict(SMALL, **_G, backbone='vit-t16', num_prompts=8, local_k=[3, 2, 2], DHW=[4, 4, 4], share_factor=2), 3, 11, 'all'),
    'gaviko_t16_full_init': (dict(FULL, **_G, backbone='vit-t16', num_prompts=32, local_k=[6, 6, 6], DHW=[10, 10, 10], share_factor=1), 2, 12, 'all'),
    # BASELINE.json configs 2/3 (the headline model) and 5
    'gaviko_b16_full_init': (dict(FULL, **_G, backbone='vit-b16', num_prompts=32, local_k=[6, 6, 6], DHW=[10, 10, 10], share_factor=1), 2, 13, 'subset'),
    'gaviko_l16_small_init': (dict(SMALL, **_G, backbone='vit-l16', num_prompts=8, local_k=[3, 2, 2], DHW=[4, 4, 4], share_factor=1), 2, 14, 'subset'),
}

VARIANT_CASES = {
    # name: (method, ctor kwargs, batch)
    'linear_t16_small': ('linear', dict(SMALL, num_classes=5, channels=1, pool='cls', backbone='vit-t16', dropout=0.0, emb_dropout=0.0), 2),
    'bitfit_t16_small': ('bitfit', dict(SMALL, num_classes=5, channels=1, pool='cls', backbone='vit-t16', dropout=0.0, emb_dropout=0.0), 2),
    'adaptformer_t16_small': ('adaptformer', dict(SMALL, num_classes=5, channels=1, pool='cls', backbone='vit-t16', dropout=0.0, emb_dropout=0.0, freeze_vit=True), 2),
    'melo_t16_small': ('melo', dict(SMALL, num_classes=5, channels=1, pool='cls', backbone='vit-t16', dropout=0.0, emb_dropout=0.0, r=4, alpha=8), 2),
    'ssf_t16_small': ('ssf', dict(SMALL, num_classes=5, channels=1, pool='cls', backbone='vit-t16', dropout=0.0, emb_dropout=0.0, freeze_vit=True), 2),
    'shallow_vpt_t16_small': ('shallow_vpt', dict(SMALL, num_classes=5, channels=1, pool='cls', backbone='vit-t16', dropout=0.0, emb_dropout=0.0,
                                                  freeze_vit=True, prompt_dropout=0.0, prompt_dim=16, num_prompts=4, deep_prompt=False), 2),
    'deep_vpt_t16_small': ('deep_vpt', dict(SMALL, num_classes=5, channels=1, pool='cls', backbone='vit-t16', dropout=0.0, emb_dropout=0.0,
                                            freeze_vit=True, prompt_dropout=0.0, prompt_dim=6, num_prompts=4, deep_prompt=True), 2),
}

# SURVEY.md §8 (f) "next" methods (f3 dvpt, f4 evp)
NEXT_CASES = {
    'dvpt_t16_small': ('dvpt', dict(SMALL, num_classes=5, channels=1, pool='mean', backbone='vit-t16', dropout=0.0, emb_dropout=0.0,
                                    num_prompts=4, freeze_vit=True), 2),
    'dvpt_cls_t16_small': ('dvpt', dict(SMALL, num_classes=5, channels=1, pool='cls', backbone='vit-t16', dropout=0.0, emb_dropout=0.0,
                                        num_prompts=6, freeze_vit=True), 2),
    # f4: EVP (model/evp.py).  scale_factor 4 is what configs/evp.yaml ships (latent width dim / 4), 32 is the constructor default
    'evp_t16_small': ('evp', dict(SMALL, num_classes=5, channels=1, pool='cls', backbone='vit-t16', dropout=0.0, emb_dropout=0.0,
                                  freeze_vit=True, scale_factor=4, input_type='fft', freq_nums=0.25, handcrafted_tune=True, embedding_tune=True), 2),
    'evp_mean_t16_small': ('evp', dict(SMALL, num_classes=5, channels=1, pool='mean', backbone='vit-t16', dropout=0.0, emb_dropout=0.0,
                                       freeze_vit=True, scale_factor=32, input_type='fft', freq_nums=0.25), 3),
}

# EVP on the reference's own seeded init at the shipped geometry and backbone (configs/evp.yaml: ViT-B, scale_factor 4 -> latent width 192, full
# 1x120x160x160 volumes): the frequency filter at H = W = 160 inside the model, the un-padded latent width, compact gradient store.
EVP_INIT_CASES = {
    'evp_b16_full_init': (dict(FULL, num_classes=5, channels=1, pool='cls', backbone='vit-b16', dropout=0.0, emb_dropout=0.0, freeze_vit=True,
                               scale_factor=4, input_type='fft', freq_nums=0.25, handcrafted_tune=True, embedding_tune=True), 2, 21, 'subset'),
}
