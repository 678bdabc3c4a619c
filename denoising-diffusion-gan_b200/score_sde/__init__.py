"""Drop-in `score_sde` package: same module paths, names and signatures as the reference's score_sde/ tree, backed by
the sm_100a kernels of ddgan_b200.  Put `denoising-diffusion-gan_b200/` ahead of the reference on sys.path and
train_ddgan.py / test_ddgan.py import these instead."""
