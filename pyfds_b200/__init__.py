"""pyfds_b200 -- B200-native drop-in for the finite-difference time-stepping path of emtpb/pyfds.

Same public names as ``pyfds`` for the hot path (``pyfds/__init__.py:1-9`` star-exports its modules):
fields, regions and material assignment, boundaries, Output probes and the acoustic / thermal models
whose ``simulate()`` runs on the CUDA step engine (``libfdsb200.so``, C ABI in ``include/fdsb200.h``).
"""

from . import fields, regions
from .acoustic_flow import *
from .acoustics import *
from .coupled_fields import *
from .coupling import *
from .regions import *
from .snapshots import *
from .thermal import *

__version__ = '0.1.0'
