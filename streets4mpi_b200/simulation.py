"""Simulation: the per-day traffic assignment (reference simulation.py:37-138) on the GPU.

Same constructor, attributes and methods as the reference class; the bodies
hand the work to libs4mpi (K1 weights -> K2 batched SSSP -> K3 walk/scatter-add,
K5 road adaptation).  ``traffic_load`` and ``cumulative_traffic_load`` stay
assignable ``array('I')`` attributes, because the reference driver replaces
them after its Allreduce (streets4mpi.py:97-100); behind the properties the
arrays live in HBM and only cross PCIe when the host actually looks at them.
``exchange()`` is the device-resident form of that Allreduce + merge.
"""
from __future__ import annotations

from array import array

import numpy as np

from . import engine
from .settings import device_settings, settings
from .tripgenerator import SampledTrips, TripTable


def _to_array_I(a):
    out = array("I")
    out.frombytes(np.ascontiguousarray(a, dtype=np.uint32).tobytes())
    return out


class Simulation(object):

    def __init__(self, street_network, trips, jam_tolerance, log_callback, device=None):
        self.street_network = street_network
        self.trips = trips
        self.jam_tolerance = jam_tolerance
        self.log_callback = log_callback
        self.step_counter = 0

        self._device = device or engine.default_device()
        for key in ("delta_factor", "block_threads", "ctas_per_sm", "pool_bytes", "batch_roots", "root_mode"):
            # S4MPI_OPTIONS (the A/B hook of engine.Device) outranks device_settings
            if device_settings.get(key) is not None and key not in self._device.env_options:
                self._device.set_option(key, device_settings[key])
        flat = street_network.flat()
        self._graph = street_network.device_graph(self._device)
        if isinstance(trips, SampledTrips):
            # drawn on the device: only the two populations cross PCIe
            self._sim = engine.DeviceSim.sampled(self._graph, trips.number_of_trips, flat.dense(trips.origins),
                                                 flat.dense(trips.goals), trips.seed, float(jam_tolerance),
                                                 engine.make_physics(settings))
            ids = flat.node_ids

            def fetch(sim=self._sim):
                o, g = sim.get_trips()
                return ids[o], ids[g]
            trips._fetch = fetch
        else:
            table = trips if isinstance(trips, TripTable) else TripTable.from_dict(trips)
            self._sim = engine.DeviceSim(self._graph, flat.dense(table.origins), flat.dense(table.goals),
                                         float(jam_tolerance), engine.make_physics(settings))
        # always: the device graph caches the limits of its first upload, the network may have moved on since
        self._sim.set_max_speed(flat.max_speed, flat.max_speed_insertion)
        self._speed_version = street_network._speed_version
        # host mirrors: None = the device copy is authoritative
        self._host_load = None
        self._host_load_dirty = False
        self._stuck_seen = 0
        self._host_cum = None            # only meaningful while _cum_on_host
        self._cum_on_host = False
        self._cum_dirty = False

    # ---------------------------------------------------------------- load arrays
    @property
    def traffic_load(self):
        if self._host_load is None:
            self._host_load = _to_array_I(self._sim.get_load())
            self._check_stuck()
        return self._host_load

    @traffic_load.setter
    def traffic_load(self, value):
        self._host_load = value
        self._host_load_dirty = True

    @property
    def cumulative_traffic_load(self):
        if not self._cum_on_host:
            cum = self._sim.get_cumulative()
            self._host_cum = None if cum is None else _to_array_I(cum)
            self._cum_on_host = True
        return self._host_cum

    @cumulative_traffic_load.setter
    def cumulative_traffic_load(self, value):
        self._host_cum = value
        self._cum_on_host = True
        self._cum_dirty = True

    def _push_host_state(self):
        if self._host_load_dirty:
            self._sim.set_load(np.frombuffer(self._host_load, dtype=np.uint32) if isinstance(self._host_load, array)
                               else np.asarray(self._host_load, dtype=np.uint32))
            self._host_load_dirty = False
        if self._cum_dirty:
            c = self._host_cum
            self._sim.set_cumulative(None if c is None else
                                     (np.frombuffer(c, dtype=np.uint32) if isinstance(c, array) else np.asarray(c, dtype=np.uint32)))
            self._cum_dirty = False
        net = self.street_network
        if net._speed_version != self._speed_version:           # change_maxspeed was called on the host object
            flat = net.flat()
            self._sim.set_max_speed(flat.max_speed, flat.max_speed_insertion)
            self._speed_version = net._speed_version

    # ---------------------------------------------------------------- the day
    def _check_stuck(self):
        """A walk that finds no predecessor (zero-weight plateau beyond the search bound) would have raised KeyError
        in the reference (simulation.py:85); here it is counted, and surfaced as soon as the host looks."""
        if not device_settings.get("warn_stuck", True):
            return
        n = self._sim.trips_stuck()
        self._stuck_seen = min(self._stuck_seen, n)          # the counters may have been reset meanwhile
        if n > self._stuck_seen:
            import warnings
            msg = "%d trips could not be walked back to their root (zero-weight plateau); their load is partial" % (n - self._stuck_seen)
            self._stuck_seen = n
            self.log_callback(msg)
            warnings.warn(msg, RuntimeWarning, stacklevel=3)

    def step(self):
        self.step_counter += 1
        self.log_callback("Preparing edges...")
        self._push_host_state()
        self._sim.step()                         # K1 -> K2 -> K3, asynchronous
        self._host_load = None
        self.log_callback("Routed", str(self._sim.n_roots), "origins on the GPU...")

    def exchange(self, group=None, virtual_ranks=None):
        """Device-resident streets4mpi.py:97-100: sum the local loads of all ranks
        (NCCL all-reduce when torch.distributed is initialised), make the total the
        new ``traffic_load`` and fold it into ``cumulative_traffic_load``.
        ``virtual_ranks``: further Simulations on this GPU that take part."""
        from . import distributed
        distributed.exchange([self] + list(virtual_ranks or []), group)

    def road_construction(self):
        self._check_stuck()                      # a synchronising call anyway (the new limits are read back)
        self._push_host_state()
        mask = self.street_network.flat().adaptable if device_settings.get("orientation_aliasing", True) else None
        self._sim.road_construction(mask)
        vis, ins = self._sim.get_max_speed()
        self.street_network._apply_max_speed(vis, ins)
        self._speed_version = self.street_network._speed_version
        self._host_cum = None
        self._cum_on_host = True
        self._cum_dirty = False

    def sync_street_network(self):
        """Copy the weights of the day back into ``street_network`` (the reference
        updated them in place, simulation.py:65)."""
        self.street_network._apply_driving_time(self._sim.weights())

    def counters(self):
        return self._sim.counters()


def calculate_driving_speed(street_length, max_speed, number_of_trips):
    """Reference simulation.py:129-138, evaluated by the device kernel (one street)."""
    return float(calculate_driving_speeds([street_length], [max_speed], [number_of_trips])[0])


def calculate_driving_speeds(street_lengths, max_speeds, numbers_of_trips):
    """Batch form of ``calculate_driving_speed`` (fp64 array in km/h)."""
    return engine.default_device().driving_speed(street_lengths, max_speeds, numbers_of_trips,
                                                 engine.make_physics(settings))
