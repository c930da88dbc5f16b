"""Host-side control flow of the sampling loop around the engine (no GPU: the engine and the sampling kernel are
replaced by recorders).  Which positions go through jk_prior_prefill, which through jk_prior_step, and which
get a token drawn - the bookkeeping of ConditionalAutoregressive2D.primed_sample (reference
prior/autoregressive.py:251-359)."""
import torch

import jukebox_b200.prior.autoregressive as ar


class FakeEngine:
    def __init__(self, capacity):
        self.prefill_capacity = capacity
        self.calls = []
        self.position = 0

    def reset(self, t0=0):
        self.position = t0

    def set_encoder_kv(self, kv):
        self.calls.append(("enc", tuple(kv.shape)))

    def prefill(self, n, P, h_out=None, **kw):
        assert self.position == 0
        self.calls.append(("prefill", n, P) if h_out is None else ("prefill", n, P, tuple(h_out.shape)))
        if h_out is not None:
            h_out.zero_()
        self.position = P

    def step(self, n, tokens=None, logits=None, **kw):
        self.calls.append(("step", self.position, logits is not None))
        if logits is not None:       # [N, bins], or the whole [N, positions, bins] preds buffer (row = position)
            (logits[:, self.position] if logits.dim() == 3 else logits).zero_()
        self.position += 1


def _model(monkeypatch, capacity):
    m = ar.ConditionalAutoregressive2D((24,), 16, width=64, depth=2, heads=1, attn_order=0, blocks=None).eval()
    eng = FakeEngine(capacity)
    monkeypatch.setattr(m, "_engine", lambda n: eng)
    monkeypatch.setattr(m.transformer, "check_cache", lambda *a, **k: None)
    drawn = []

    def fake_sample(logits, temp, seed, position, tokens):
        drawn.append(position)
        tokens[:, position] = position % 16
    monkeypatch.setattr(ar, "sample_categorical", fake_sample)
    return m, eng, drawn


def test_primed_sample_prefills_the_given_tokens_once(monkeypatch):
    m, eng, drawn = _model(monkeypatch, capacity=512)
    prime = torch.randint(0, 16, (3, 7))
    z = m.primed_sample(3, prime, fp16=True, temp=0.9, sample_tokens=12)
    assert eng.calls[0] == ("prefill", 3, 7)
    steps = [c for c in eng.calls if c[0] == "step"]
    assert [c[1] for c in steps] == list(range(7, 12)) and all(c[2] for c in steps)   # every later step wants logits
    assert drawn == list(range(7, 12))
    assert torch.equal(z[:, :7], prime) and z.shape == (3, 12)


def test_get_preds_takes_the_given_positions_logits_from_the_prefill(monkeypatch):
    import jukebox_b200.transformer.f32 as f32
    m, eng, drawn = _model(monkeypatch, capacity=512)
    monkeypatch.setattr(f32, "linear_nk", lambda x, w: torch.full((x.shape[0], w.shape[0]), 7.0))
    prime = torch.randint(0, 16, (2, 7))
    z, preds = m.primed_sample(2, prime, fp16=True, get_preds=True, sample_tokens=10)
    assert eng.calls[0] == ("prefill", 2, 7, (2, 7, 64))                 # activations of the given positions requested
    assert [c[:2] for c in eng.calls[1:]] == [("step", 7), ("step", 8), ("step", 9)]
    assert preds.shape == (2, 10, 16) and bool((preds[:, :7] == 7.0).all()) and bool((preds[:, 7:] == 0.0).all())
    assert drawn == [7, 8, 9] and torch.equal(z[:, :7], prime)


def test_stepping_when_prefill_is_unavailable(monkeypatch):
    for capacity, get_preds in ((0, False), (4, False), (0, True)):
        m, eng, drawn = _model(monkeypatch, capacity)
        prime = torch.randint(0, 16, (2, 7))
        out = m.primed_sample(2, prime, fp16=True, get_preds=get_preds, sample_tokens=10)
        assert all(c[0] == "step" for c in eng.calls)
        assert [c[1] for c in eng.calls] == list(range(10))
        # logits are only requested where something is done with them
        assert [c[2] for c in eng.calls] == [get_preds or t >= 7 for t in range(10)]
        assert drawn == [7, 8, 9]
        if get_preds:
            assert out[1].shape == (2, 10, 16)


def test_ancestral_sampling_never_prefills(monkeypatch):
    m, eng, drawn = _model(monkeypatch, capacity=512)
    z = m.sample(2, fp16=True, sample_tokens=5)
    assert [c[0] for c in eng.calls] == ["step"] * 5 and drawn == list(range(5)) and z.shape == (2, 5)


def test_logit_bias_is_computed_once_per_window_and_passed_to_every_step(monkeypatch):
    """a prior that adds x_cond behind the stack (autoregressive.py:226-227): x_cond . x_out^T once per window, every step
    gets it (jkb200.h: jk_step_args.logit_bias) - only when the engine multiplies the logits on the tensor cores"""
    import jukebox_b200.transformer.f32 as f32
    for has_gemm in (True, False):
        m = ar.ConditionalAutoregressive2D((24,), 16, width=64, depth=2, heads=1, attn_order=0, blocks=None, x_cond=True).eval()
        eng = FakeEngine(512)
        eng.has_logits_gemm = has_gemm
        seen = []
        orig_step = eng.step

        def step(n, logit_bias=None, **kw):
            seen.append(logit_bias)
            return orig_step(n, **kw)
        eng.step = step
        monkeypatch.setattr(m, "_engine", lambda n: eng)
        monkeypatch.setattr(m.transformer, "check_cache", lambda *a, **k: None)
        monkeypatch.setattr(ar, "sample_categorical", lambda logits, temp, seed, position, tokens: None)
        calls = []

        def fake_linear(x, w):
            calls.append((tuple(x.shape), tuple(w.shape)))
            return torch.ones(x.shape[0], w.shape[0])
        monkeypatch.setattr(f32, "linear_nk", fake_linear)
        xc = torch.randn(2, 24, 64)
        m.sample(2, x_cond=xc, fp16=True, sample_tokens=6)
        assert m.add_cond_after_transformer
        if has_gemm:
            assert calls == [((2 * 24, 64), (16, 64))]                       # one GEMM over all positions of the window
            assert len(seen) == 6 and all(b is not None and tuple(b.shape) == (2, 24, 16) for b in seen)
            assert all(b is seen[0] for b in seen)
        else:
            assert calls == [] and seen == [None] * 6
