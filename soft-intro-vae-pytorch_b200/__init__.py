"""B200-native Soft-IntroVAE training engine (drop-in for soft_intro_vae/train_soft_intro_vae.py of
taldatech/soft-intro-vae-pytorch).  The directory name is not a valid identifier; import it with
``importlib.import_module("soft-intro-vae-pytorch_b200")`` or through the root-level alias ``sivae_b200``.

Sub-modules
  lib                             ctypes binding of csrc/ -> libsivae_b200.so (C ABI: include/sivae.h)
  engine                          torch-side owner of the flat parameter / workspace buffers handed to the library
  train_soft_intro_vae            mirror of the reference module of the same name (classes, helpers, train fn)
  train_soft_intro_vae_bootstrap  bootstrap (target decoder) twin
  train_soft_intro_vae_2d         2-D toy twin (CPU plumbing config)
"""
