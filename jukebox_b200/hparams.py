"""Hyper-parameter registry: same names, same resolved values and the same `setup_hparams`
semantics as the reference (jukebox/hparams.py) - tests/test_host_cpu.py compares every resolved set
against a dump of the reference's registry (tests/golden/hparams.json).
"""


class Hyperparams(dict):
    """dict with attribute access"""

    def __getattr__(self, attr):
        return self[attr]

    def __setattr__(self, attr, value):
        self[attr] = value


HPARAMS_REGISTRY = {}
DEFAULTS = {}
REMOTE_PREFIX = 'https://openaipublic.azureedge.net/'


def setup_hparams(hparam_set_names, kwargs):
    """all DEFAULTS groups, then each named set in order, then kwargs; unknown keys are an error"""
    H = Hyperparams()
    if not isinstance(hparam_set_names, tuple):
        hparam_set_names = hparam_set_names.split(",")
    hparam_sets = [HPARAMS_REGISTRY[x.strip()] for x in hparam_set_names if x] + [kwargs]
    for group in DEFAULTS.values():
        H.update(group)
    for hps in hparam_sets:
        for k in hps:
            if k not in H:
                raise ValueError(f"{k} not in default args")
        H.update(**hps)
    H.update(**kwargs)
    return H


def _register(name, *bases, **values):
    h = Hyperparams(**values)
    for b in bases:
        h.update(b)
    HPARAMS_REGISTRY[name] = h
    return h


_register("teeny")
_register("easy", sr=22050)

# ---- released models -----------------------------------------------------------------------
_register("vqvae", levels=3, downs_t=(3, 2, 2), strides_t=(2, 2, 2), emb_width=64, l_bins=2048, l_mu=0.99,
          commit=0.02, spectral=0.0, multispectral=1.0, hvqvae_multipliers=(2, 1, 1), loss_fn='lmix', lmix_l2=1.0,
          lmix_linf=0.02, width=32, depth=4, m_conv=1.0, dilation_growth_rate=3,
          restore_vqvae=REMOTE_PREFIX + 'jukebox/models/5b/vqvae.pth.tar')

_labels_v2 = dict(y_bins=(120, 4111), t_bins=128, max_bow_genre_size=5, n_vocab=80)
_labels_v3 = dict(y_bins=(604, 7898), t_bins=64, max_bow_genre_size=1, n_vocab=79)

_upsamplers = dict(n_ctx=8192, prior_width=1920, prior_depth=72, heads=1, attn_order=2, blocks=128, init_scale=0.4,
                   c_res=1, cond_width=1024, cond_depth=16, cond_dilation_growth_rate=3, cond_dilation_cycle=8,
                   cond_c_res=1, use_tokens=False, prime_loss_fraction=0.0, fp16_params=False, **_labels_v2)
_register("upsampler_level_0", _upsamplers, level=0,
          restore_prior=REMOTE_PREFIX + 'jukebox/models/5b/prior_level_0.pth.tar')
_register("upsampler_level_1", _upsamplers, level=1, cond_res_scale=True,
          restore_prior=REMOTE_PREFIX + 'jukebox/models/5b/prior_level_1.pth.tar')

_register("prior_5b", _labels_v2, level=2, n_ctx=8192, prior_width=4800, prior_depth=72, heads=8, attn_order=2,
          blocks=128, init_scale=0.1, c_res=1, beta2=0.925, min_duration=60.0, max_duration=600.0, use_tokens=False,
          n_tokens=0, prime_loss_fraction=0.0, merged_decoder=True, fp16_params=True,
          restore_prior=REMOTE_PREFIX + 'jukebox/models/5b/prior_level_2.pth.tar')

_register("prior_5b_lyrics", _labels_v2, level=2, n_ctx=8192, prior_width=4800, prior_depth=79, heads=8,
          attn_order=10, blocks=128, init_scale=0.1, c_res=1, prime_width=1280, prime_depth=18, prime_heads=4,
          prime_attn_order=2, prime_blocks=32, prime_init_scale=0.7, prime_c_res=1, min_duration=23.8,
          max_duration=600.0, use_tokens=True, n_tokens=512, prime_loss_fraction=0.4, merged_decoder=True,
          fp16_params=True, alignment_layer=68, alignment_head=2,
          restore_prior=REMOTE_PREFIX + 'jukebox/models/5b_lyrics/prior_level_2.pth.tar')

_register("prior_1b_lyrics", _labels_v3, level=2, n_ctx=6144, prior_width=2048, prior_depth=72, heads=2,
          attn_order=12, blocks=64, init_scale=0.2, c_res=1, labels_v3=True, min_duration=17.84, max_duration=600.0,
          use_tokens=True, n_tokens=384, prime_loss_fraction=0.4, single_enc_dec=True, fp16_params=False,
          alignment_layer=63, alignment_head=0,
          restore_prior=REMOTE_PREFIX + 'jukebox/models/1b_lyrics/prior_level_2.pth.tar')

# ---- small models (README training recipes) --------------------------------------------------
_register("small_vqvae", sr=22050, levels=2, downs_t=(5, 3), strides_t=(2, 2), emb_width=64, l_bins=1024,
          l_mu=0.99, commit=0.02, spectral=0.0, multispectral=1.0, loss_fn='l2', width=32, depth=4, m_conv=1.0,
          dilation_growth_rate=3)
_small_prior = _register("small_prior", n_ctx=8192, prior_width=1024, prior_depth=48, heads=1, c_res=1,
                         attn_order=2, blocks=64, init_scale=0.7)
_small_labels = dict(labels=True, labels_v3=True, y_bins=(10, 100), max_bow_genre_size=1, min_duration=60.0,
                     max_duration=600.0, t_bins=64)
_register("small_labelled_prior", _small_prior, **_small_labels)
_register("small_single_enc_dec_prior", _small_labels, n_ctx=6144, prior_width=1024, prior_depth=48, heads=2,
          attn_order=12, blocks=64, init_scale=0.7, c_res=1, prime_loss_fraction=0.4, single_enc_dec=True,
          use_tokens=True, n_tokens=384, n_vocab=79)
_register("small_sep_enc_dec_prior", _small_labels, n_ctx=6144, prior_width=1024, prior_depth=50, heads=2,
          attn_order=8, blocks=64, init_scale=0.7, c_res=1, prime_width=256, prime_depth=9, prime_heads=2,
          prime_attn_order=2, prime_blocks=32, prime_init_scale=0.7, prime_c_res=1, prime_loss_fraction=0.4,
          use_tokens=True, n_tokens=384, n_vocab=79)
_register("small_upsampler", n_ctx=8192, prior_width=1024, prior_depth=48, heads=1, c_res=1, attn_order=2,
          blocks=64, init_scale=0.7, cond_width=512, cond_depth=16, cond_dilation_growth_rate=3,
          cond_dilation_cycle=8, cond_c_res=1)
_register("all_fp16", fp16=True, fp16_params=True, fp16_opt=True, fp16_scale_window=250)
_register("cpu_ema", ema=True, cpu_ema=True, cpu_ema_freq=100, ema_fused=False)

# ---- defaults: every key a set or a CLI override may touch -----------------------------------
_D = {
    "rcall": dict(rcall_command="<unknown_rcall_command>", git_commit="<unknown_git_commit>"),
    "script": dict(name='', debug_mem=False, debug_eval_files=False, debug_speed=False, debug_iters=100,
                   debug_batch=False, debug_grad_accum=False, debug_inputs=False, local_path='',
                   local_logdir='logs', max_len=24, max_log=32, save=True, save_iters=20000, seed=0, prior=False,
                   log_steps=100, func=''),
    "data": dict(audio_files_dir='', finetune='', english_only=False, bs=1, bs_sample=1, nworkers=1, aug_shift=False,
                 aug_blend=False, train_test_split=0.9, train_shrink_factor=1.0, test_shrink_factor=1.0, p_unk=0.1,
                 min_duration=None, max_duration=None, n_tokens=0, n_vocab=0, use_tokens=False, curr_epoch=-1),
    "vqvae": dict(restore_vqvae='', levels=2, downs_t=(1, 1), strides_t=(2, 2), hvqvae_multipliers=None,
                  revival_threshold=1.0, emb_width=64, l_bins=512, l_mu=0.99, commit=1.0, spectral=0.0,
                  multispectral=1.0, loss_fn='l2', linf_k=2048, lmix_l1=0.0, lmix_l2=0.0, lmix_linf=0.0,
                  use_bottleneck=True),
    "vqvae_conv_block": dict(depth=3, width=128, m_conv=1.0, dilation_growth_rate=1, dilation_cycle=None,
                             vqvae_reverse_decoder_dilation=True),
    "prior": dict(restore_prior='', restore_prior_ddp=False, max_bow_genre_size=None, y_bins=0, level=0,
                  cond_levels=None, t_bins=64, y_cond_as_bias=False, copy_input=False, merged_decoder=False,
                  single_enc_dec=False, alignment_layer=None, alignment_head=None),
    "prior_attn_block": dict(n_ctx=1024, prior_depth=3, prior_width=128, heads=1, attn_order=0, blocks=None,
                             spread=None, attn_dropout=0.0, resid_dropout=0.0, emb_dropout=0.0, zero_out=False,
                             res_scale=False, pos_init=False, init_scale=1.0, m_attn=0.25, m_mlp=1.0, c_res=0,
                             c_attn=0, c_mlp=0),
    "cond_conv_block": dict(cond_depth=3, cond_width=128, cond_m_conv=1.0, cond_zero_out=False,
                            cond_res_scale=False, cond_dilation_growth_rate=1, cond_dilation_cycle=None,
                            cond_c_res=0),
    "sample": dict(primed_chunk_size=None, selected_artists='', temp_top=1.0, temp_rest=0.99,
                   sample_length_in_seconds=24, total_sample_length_in_seconds=240),
    "prime": dict(prime_loss_fraction=0.1, restore_decoder=''),
    "prime_attn_block": dict(prime_depth=3, prime_width=128, prime_heads=1, prime_attn_order=0, prime_blocks=None,
                             prime_spread=None, prime_attn_dropout=0.0, prime_resid_dropout=0.0,
                             prime_emb_dropout=0.0, prime_zero_out=False, prime_res_scale=False,
                             prime_pos_init=False, prime_init_scale=1.0, prime_m_attn=0.25, prime_m_mlp=1.0,
                             prime_c_res=0, prime_c_attn=0, prime_c_mlp=0, prime_rel_attn=False,
                             prime_posemb_timescale=10000),
    "opt": dict(epochs=10000, lr=0.0003, clip=1.0, beta1=0.9, beta2=0.999, ignore_grad_norm=0, weight_decay=0.0,
                eps=1e-08, lr_warmup=100.0, lr_decay=10000000000.0, lr_gamma=1.0, lr_scale=1.0,
                lr_use_linear_decay=False, lr_start_linear_decay=0, lr_use_cosine_decay=False),
    "fp16": dict(fp16=False, fp16_params=False, fp16_loss_scale=None, fp16_scale_window=1000.0, fp16_opt=False),
    "train_test_eval": dict(labels=True, labels_v3=False, dump=False, ema=True, ema_fused=True, cpu_ema=False,
                            cpu_ema_freq=100, reset_best_loss=False, reset_step=False, reset_opt=False,
                            reset_shd=False, train=False, test=False, sample=False, sampler='ancestral',
                            codes_logdir='', date=None, labeller='top_genres', label_line=0,
                            iters_before_update=1, grad_accum_iters=0, mu=None, piped=False, pipe_depth=8,
                            break_train=1e10, break_test=1e10, exit_train=1e10),
    "audio": dict(n_fft=1024, hop_length=256, window_size=1024, sr=44100, channels=2, wav='', n_inps=1, n_hops=2,
                  n_segment=1, n_total_segment=1, n_segment_each=1, prime_chunks=4, sample_length=0,
                  sample_hop_length=30000, max_silence_pad_length=0, ignore_boundaries=False,
                  use_nonrelative_specloss=True, multispec_loss_n_fft=(2048, 1024, 512),
                  multispec_loss_hop_length=(240, 120, 50), multispec_loss_window_size=(1200, 600, 240)),
    "distributed": dict(bucket=128),
}
for _k, _v in _D.items():
    DEFAULTS[_k] = Hyperparams(**_v)
