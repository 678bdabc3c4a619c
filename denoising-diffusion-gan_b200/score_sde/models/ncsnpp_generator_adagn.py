"""Drop-in for score_sde/models/ncsnpp_generator_adagn.py: `NCSNpp(config).forward(x, time_cond, z)`."""
from ddgan_b200.modules import NCSNpp as _NCSNpp
from . import utils


@utils.register_model(name='ncsnpp')
class NCSNpp(_NCSNpp):
    pass
