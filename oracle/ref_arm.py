"""
ORACLE TOOLING -- TEST INFRASTRUCTURE ONLY (bench.py's reference arm / cpu_baseline leg, tests/).

Drives the UNMODIFIED reference trainer (`oracle/_ref/...`, produced by oracle/build_ref.py; /root/reference when present) on
the host CPU and times its introspective iterations.  Nothing of the reference is edited: the module is imported as it is and
only module GLOBALS are substituted, the way SURVEY App. D describes -- `matplotlib` (absent in this image) by a stub, the
dataset class by a synthetic in-memory data set of the workload's shape (no files, no network), and `tqdm` by an iterator
that takes a timestamp every time the training loop asks for its next batch.  The time between two such requests is one full
iteration of the reference's own loop body (:542-646: both halves, both optimiser steps, the `.item()` logging).
"""
import importlib
import os
import sys
import tempfile
import time
from unittest.mock import MagicMock

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_VARIANTS = {False: ("soft_intro_vae", "train_soft_intro_vae"), True: ("soft_intro_vae_bootstrap", "train_soft_intro_vae_bootstrap")}
_DATASET = {32: "cifar10", 128: "celeb128", 256: "celeb256"}


def reference_root():
    """directory holding the reference's script directories, or None"""
    for cand in (os.path.join(ROOT, "oracle", "_ref"), os.environ.get("SIVAE_REFERENCE_ROOT", "/root/reference")):
        if cand and os.path.isfile(os.path.join(cand, "soft_intro_vae", "train_soft_intro_vae.py")):
            return cand
    return None


def import_reference(bootstrap=False):
    """the reference trainer module, imported unmodified (matplotlib stubbed: the image has none)"""
    root = reference_root()
    if root is None:
        raise FileNotFoundError("no reference scripts: run `python oracle/build_ref.py` where /root/reference exists")
    subdir, modname = _VARIANTS[bool(bootstrap)]
    if "matplotlib" not in sys.modules:
        try:
            importlib.import_module("matplotlib")
        except ImportError:
            mpl, plt = MagicMock(), MagicMock()
            plt.subplots = lambda *a, **k: (MagicMock(), MagicMock())
            mpl.pyplot = plt
            sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
    for m in ("dataset", "metrics", "metrics.fid_score", "metrics.inception", modname):
        sys.modules.pop(m, None)
    sys.path.insert(0, os.path.join(root, subdir))
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")            # dataset.py: `is not 'RGB'` SyntaxWarnings
            mod = importlib.import_module(modname)
    finally:
        sys.path.pop(0)
        for m in ("dataset", modname):                 # do not leave the reference's modules importable by name
            sys.modules.pop(m, None)
    return mod


class _Synth(torch.utils.data.Dataset):
    def __init__(self, n, size, labelled, seed=1234):
        self.x = torch.rand(n, 3, size, size, generator=torch.Generator().manual_seed(seed))
        self.labelled = labelled

    def __len__(self):
        return self.x.size(0)

    def __getitem__(self, i):
        return (self.x[i], 0) if self.labelled else self.x[i]


def time_reference_iterations(size, zdim, batch, beta_neg, bootstrap=False, warmup=1, steps=1, threads=None, seed=-1, device="cpu"):
    """run the reference's own train function for (warmup + steps) iterations of `batch` images on `device` ("cpu": the CPU arm;
    "cuda:N": the unmodified reference through stock PyTorch / cuDNN on that GPU -- its loop reads ten scalars with .item() every
    iteration (:628-639), so the host time between two batch requests is the device time of the iteration); returns
    (seconds per timed iteration, list of all iteration times).  seed = -1 like the reference's default (a fixed seed would also
    switch cudnn.deterministic on, :372)."""
    ref = import_reference(bootstrap)
    threads = threads or (os.cpu_count() or 1)
    torch.set_num_threads(threads)
    n_it = warmup + steps
    stamps = []

    class Bar:                                           # stands in for tqdm(iterable=loader) (:506)
        def __init__(self, iterable=None, **k):
            self.it = iterable

        def __iter__(self):
            it = iter(self.it)
            while True:
                stamps.append(time.perf_counter())       # the loop asks for its next batch: previous iteration is complete
                try:
                    b = next(it)
                except StopIteration:
                    return
                yield b

        def set_description_str(self, *a, **k):
            pass

        def set_postfix(self, **k):
            pass

        def close(self):
            pass

    ref.tqdm = Bar
    ds = _Synth(n_it * batch, size, labelled=size == 32)
    if size == 32:
        ref.CIFAR10 = lambda *a, **k: ds                 # looked up as a module global at :379
    else:
        ref.ImageDatasetFromFile = lambda *a, **k: ds    # :388-392 / :400-404
    fn = getattr(ref, "train_soft_intro_vae")
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp(prefix="sivae_refarm_")
    run_dir = os.path.join(tmp, "run")
    os.makedirs(run_dir)
    if size != 32:                                       # the file-list code of the dataset switch wants image NAMES to exist
        d = os.path.join(tmp, "data", "celeb256", "img_align_celeba")
        os.makedirs(d)
        for i in range(4):
            open(os.path.join(d, "%06d.jpg" % i), "w").close()
    os.chdir(run_dir)
    stdout = sys.stdout
    sys.stdout = open(os.devnull, "w")
    try:
        kw = dict(dataset=_DATASET[size], z_dim=zdim, batch_size=batch, num_workers=0, num_epochs=1, num_vae=0, beta_kl=1.0,
                  beta_neg=beta_neg, beta_rec=1.0, device=torch.device(device), seed=seed, test_iter=10 ** 9, save_interval=50,
                  start_epoch=0, lr_e=2e-4, lr_d=2e-4)
        if bootstrap:
            kw.update(gamma_r=1.0, copy_to_target_freq=1)
        fn(**kw)
    finally:
        sys.stdout.close()
        sys.stdout = stdout
        os.chdir(cwd)
        import shutil
        shutil.rmtree(tmp, ignore_errors=True)
    if torch.device(device).type == "cuda":
        torch.cuda.synchronize(torch.device(device))
        torch.cuda.empty_cache()
    its = [b - a for a, b in zip(stamps[:-1], stamps[1:])]
    assert len(its) == n_it, (len(its), n_it)
    timed = its[warmup:]
    return sum(timed) / len(timed), its


def run_reference_training(n_images, batch, zdim, beta_neg, seed, device="cpu", size=32, bootstrap=False, threads=None):
    """the UNMODIFIED reference `train_soft_intro_vae()` for one epoch over a synthetic data set (the tensor the drop-in's
    `synthetic<size>:<n>` data set holds: torch.rand with generator seed 1234) with a fixed seed; returns the tqdm postfix
    dictionaries of its iterations (:629-631: r_loss, kl, diff_kl, expelbo_f) -- the trainer-level trace the drop-in is held to"""
    ref = import_reference(bootstrap)
    if threads:
        torch.set_num_threads(threads)
    trace = []

    class Bar:
        def __init__(self, iterable=None, **k):
            self.it = iterable

        def __iter__(self):
            return iter(self.it)

        def set_description_str(self, *a, **k):
            pass

        def set_postfix(self, **k):
            trace.append({n: float(v) for n, v in k.items()})

        def close(self):
            pass

    ref.tqdm = Bar
    ds = _Synth(n_images, size, labelled=size == 32)
    if size == 32:
        ref.CIFAR10 = lambda *a, **k: ds
    else:
        ref.ImageDatasetFromFile = lambda *a, **k: ds
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp(prefix="sivae_reftrace_")
    run_dir = os.path.join(tmp, "run")
    os.makedirs(run_dir)
    if size != 32:
        d = os.path.join(tmp, "data", "celeb256", "img_align_celeba")
        os.makedirs(d)
        for i in range(4):
            open(os.path.join(d, "%06d.jpg" % i), "w").close()
    os.chdir(run_dir)
    stdout = sys.stdout
    sys.stdout = open(os.devnull, "w")
    try:
        kw = dict(dataset=_DATASET[size], z_dim=zdim, batch_size=batch, num_workers=0, num_epochs=1, num_vae=0, beta_kl=1.0,
                  beta_neg=beta_neg, beta_rec=1.0, device=torch.device(device), seed=seed, test_iter=1000, save_interval=50,
                  start_epoch=0, lr_e=2e-4, lr_d=2e-4)
        if bootstrap:
            kw.update(gamma_r=1.0, copy_to_target_freq=1)
        getattr(ref, "train_soft_intro_vae")(**kw)
    finally:
        sys.stdout.close()
        sys.stdout = stdout
        os.chdir(cwd)
        import shutil
        shutil.rmtree(tmp, ignore_errors=True)
    return trace
