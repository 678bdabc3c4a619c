"""Drop-in for score_sde/models/discriminator.py: Discriminator_small / Discriminator_large `.forward(x, t, x_t)`."""
from ddgan_b200.modules import Discriminator_large, Discriminator_small  # noqa: F401
