"""Checkpoint restore through the reference's entry points (make_models.py:17-61): a file in the released
format ({'model': state_dict with optional 'module.' prefixes, 'step': n}) loads strictly by parameter name."""
import torch

from jukebox_b200.hparams import setup_hparams
from jukebox_b200.make_models import make_vqvae, make_prior


def _ckpt(model, path, prefix="", step=1234):
    fresh = {}                                           # tied parameters (x_out.weight is x_emb.weight) keep one value
    sd = {}
    for k, v in model.state_dict().items():
        key = v.data_ptr()
        if key not in fresh:
            fresh[key] = torch.randn_like(v) if v.is_floating_point() else v.clone()
        sd[prefix + k] = fresh[key]
    torch.save({"model": sd, "step": step}, path)
    return sd


def test_vqvae_and_prior_restore_by_name(tmp_path):
    vq0 = make_vqvae(setup_hparams("small_vqvae", dict(sample_length=8192, restore_vqvae="")), "cpu")
    p = str(tmp_path / "vqvae.pth.tar")
    sd = _ckpt(vq0, p, prefix="module.")                 # DDP-saved checkpoints carry the 'module.' prefix
    vq = make_vqvae(setup_hparams("small_vqvae", dict(sample_length=8192, restore_vqvae=p)), "cpu")
    assert vq.step == 1234
    for k, v in vq.state_dict().items():
        assert torch.equal(v, sd["module." + k]), k
    assert not any(q.requires_grad for q in vq.parameters())      # frozen, eval mode (make_models.py:96-98)

    hps = setup_hparams("small_vqvae,small_prior", dict(sample_length=8192 * 8, restore_vqvae="", restore_prior="", level=1,
                                                        levels=2, labels=False))
    pr0 = make_prior(hps, vq0, "cpu")
    p2 = str(tmp_path / "prior.pth.tar")
    sd2 = _ckpt(pr0, p2, step=7)
    hps2 = setup_hparams("small_vqvae,small_prior", dict(sample_length=8192 * 8, restore_vqvae="", restore_prior=p2, level=1,
                                                         levels=2, labels=False))
    pr = make_prior(hps2, vq0, "cpu")
    assert pr.step == 7
    own = {k: v for k, v in pr.state_dict().items()}
    assert set(own) == set(sd2)
    for k, v in own.items():
        assert torch.equal(v.float(), sd2[k].float()), k


def test_restore_rejects_unknown_names(tmp_path):
    vq0 = make_vqvae(setup_hparams("small_vqvae", dict(sample_length=8192, restore_vqvae="")), "cpu")
    sd = dict(vq0.state_dict())
    sd["decoders.0.not_a_parameter"] = torch.zeros(1)
    p = str(tmp_path / "bad.pth.tar")
    torch.save({"model": sd}, p)
    try:
        make_vqvae(setup_hparams("small_vqvae", dict(sample_length=8192, restore_vqvae=p)), "cpu")
    except RuntimeError as e:
        assert "not_a_parameter" in str(e)
    else:
        raise AssertionError("strict load must fail on unexpected keys")


def test_fp16_params_are_converted_before_restore(tmp_path):
    """make_models.py:174-179: `_convert_conv_weights_to_fp16` runs BEFORE restore_model, so an fp32 checkpoint is
    cast into fp16 Conv1D weights by load_state_dict while biases / LayerNorm / embeddings stay fp32."""
    vq0 = make_vqvae(setup_hparams("small_vqvae", dict(sample_length=8192, restore_vqvae="")), "cpu")
    over = dict(sample_length=8192 * 8, restore_vqvae="", level=1, levels=2, labels=False, prior_depth=3, prior_width=64)
    pr0 = make_prior(setup_hparams("small_vqvae,small_prior", dict(restore_prior="", **over)), vq0, "cpu")
    p = str(tmp_path / "prior32.pth.tar")
    sd = _ckpt(pr0, p)
    assert all(v.dtype == torch.float32 for v in sd.values() if v.is_floating_point())
    pr = make_prior(setup_hparams("small_vqvae,small_prior,all_fp16", dict(restore_prior=p, **over)), vq0, "cpu")
    for k, v in pr.state_dict().items():
        conv_w = k.endswith((".c_attn.w", ".c_proj.w", ".c_fc.w", ".c_enc_kv.w"))
        assert v.dtype == (torch.float16 if conv_w else torch.float32), (k, v.dtype)
        assert torch.equal(v.float(), sd[k].to(v.dtype).float()), k        # one rounding, at load time


def test_sampling_resume_file_round_trip(tmp_path, monkeypatch):
    """the per-level `data.pth.tar` of sample.py:116 ({zs, labels, sampling_kwargs, x}) written by _sample and
    read back by load_codes, incl. the cut to a shorter duration (sample.py:164-175)"""
    from jukebox_b200.hparams import Hyperparams
    from jukebox_b200.sample import save_level, load_codes
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)      # no GPU here: load_codes ends in .cuda()

    class P:
        def __init__(self, r):
            self.raw_to_tokens = r
    priors = [P(8), P(32), P(128)]
    n = 3
    zs = [torch.randint(0, 2048, (n, 1024 // p.raw_to_tokens * 16)) for p in priors]
    labels = [dict(y=torch.arange(n * 5).view(n, 5), info=[dict(artist="a")] * n) for _ in priors]
    kw = [dict(temp=0.99, fp16=True, max_batch_size=16, chunk_size=32) for _ in priors]
    x = torch.randn(n, 16384, 1)
    logdir = save_level(str(tmp_path / "run"), 2, zs, labels, kw, x)
    data = torch.load(logdir + "/data.pth.tar", weights_only=False)
    assert set(data) == {"zs", "labels", "sampling_kwargs", "x"}
    assert all(torch.equal(a, b) for a, b in zip(data["zs"], zs)) and torch.equal(data["x"], x)
    hps = Hyperparams(n_samples=n)
    back = load_codes(logdir + "/data.pth.tar", None, priors, hps)
    assert all(torch.equal(a, b) for a, b in zip(back, zs))
    cut = load_codes(logdir + "/data.pth.tar", 128 * 40, priors, hps)
    assert [z.shape[1] for z in cut] == [128 * 40 // p.raw_to_tokens for p in priors]
    assert all(torch.equal(c, z[:, :c.shape[1]]) for c, z in zip(cut, zs))
