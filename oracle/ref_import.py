"""Import the UNMODIFIED reference (/root/reference/jukebox) on a CPU-only box.

TEST INFRASTRUCTURE ONLY.  This file exists so that `oracle/make_golden.py`
can run the real reference here (build container, no GPU) and write golden
vectors under tests/golden/.  Nothing in the product (`jukebox_b200/`) may
import it, and it is never used on the GPU box (the reference tree does not
exist there).

What has to be shimmed:
  * five third-party modules the reference imports but never uses on the
    arithmetic path: fire, soundfile, librosa, unidecode, av
  * a 1-rank gloo process group (jukebox/utils/dist_adapter.py:21-25 calls
    torch.distributed.get_rank unconditionally)
  * hard-coded CUDA placement (`.cuda()`, `device='cuda'`, t.cuda.LongTensor)
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("JUKEBOX_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "jukebox"))


_done = False


def load_reference():
    """Returns the imported `jukebox` reference package (CPU-shimmed)."""
    global _done
    import torch
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if not _done:
        sys.dont_write_bytecode = True
        for name in ("fire", "soundfile", "librosa", "unidecode", "av"):
            if name not in sys.modules:
                sys.modules[name] = types.ModuleType(name)
        sys.modules["unidecode"].unidecode = lambda s: s
        sys.modules["fire"].Fire = lambda *a, **k: None
        if REFERENCE_ROOT not in sys.path:
            sys.path.insert(0, REFERENCE_ROOT)
        import torch.distributed as dist
        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            if "MASTER_PORT" not in os.environ:      # any free port: two importers may run side by side
                import socket
                with socket.socket() as sk:
                    sk.bind(("127.0.0.1", 0))
                    os.environ["MASTER_PORT"] = str(sk.getsockname()[1])
            dist.init_process_group("gloo", rank=0, world_size=1)
        if not torch.cuda.is_available():
            torch.Tensor.cuda = lambda s, *a, **k: s
            torch.nn.Module.cuda = lambda s, *a, **k: s
            torch.cuda.LongTensor = torch.LongTensor
            torch.cuda.empty_cache = lambda: None

            def _cpuify(fn):
                def wrapped(*a, **k):
                    dev = k.get("device", None)
                    if isinstance(dev, str) and dev.startswith("cuda"):
                        k["device"] = "cpu"
                    return fn(*a, **k)
                return wrapped
            for fname in ("zeros", "arange", "tensor", "ones", "empty", "randint"):
                setattr(torch, fname, _cpuify(getattr(torch, fname)))
        _done = True
    import jukebox  # noqa: F401  (the reference package)
    return sys.modules["jukebox"]
