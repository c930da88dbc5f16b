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
