"""streets4mpi_b200 -- B200-native traffic-assignment hot path of Streets4MPI.

Host side: the reference's Python API (``Streets4MPI`` / ``Simulation`` /
``StreetNetwork`` / ``TripGenerator``, ``settings``, ``persistence``, ``utils``)
over ``libs4mpi.so`` (hand-written sm_100a CUDA behind the C-ABI of
``include/s4mpi.h``).  No CPU fallback: importing works anywhere, computing
needs the built library and a CUDA device.
"""
from .settings import settings  # noqa: F401
from .utils import merge_arrays  # noqa: F401
from .streetnetwork import StreetNetwork  # noqa: F401
from .tripgenerator import SampledTrips, TripGenerator, TripTable  # noqa: F401
from .simulation import Simulation, calculate_driving_speed  # noqa: F401

__all__ = ["settings", "merge_arrays", "StreetNetwork", "TripGenerator", "TripTable", "SampledTrips", "Simulation",
           "calculate_driving_speed"]
