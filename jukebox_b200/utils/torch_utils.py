import gc

import torch as t


def freeze_model(model):
    model.eval()
    for p in model.parameters():
        p.requires_grad = False


def empty_cache():
    gc.collect()
    if t.cuda.is_available():
        t.cuda.empty_cache()


def assert_shape(x, exp_shape):
    assert tuple(x.shape) == tuple(exp_shape), f"Expected {tuple(exp_shape)} got {tuple(x.shape)}"
