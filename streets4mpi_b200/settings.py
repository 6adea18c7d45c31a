"""Configuration dict, key for key the reference's (settings.py:24-44).

``device_settings`` holds the knobs that only exist on the GPU path; the
reference dict itself is unchanged so existing settings files keep working.
"""

settings = {
    # technical settings
    "osm_file": "osm/test.osm",
    "logging": "stdout",
    "persist_traffic_load": True,
    "random_seed": 3756917,

    # simulation settings
    "max_simulation_steps": 10,
    "number_of_residents": 100,
    "use_residential_origins": False,
    "traffic_period_duration": 8,      # h   (only read by the reference's dead variant, simulation.py:122)
    "car_length": 4,                   # m
    "min_breaking_distance": 0.001,    # m
    "braking_deceleration": 7.5,       # m/s^2
    "steps_between_street_construction": 10,
    "trip_volume": 1,
}

device_settings = {
    # S4_ROOT_ORIGIN (0): one tree per distinct origin, the reference's grouping (simulation.py:71).
    # S4_ROOT_GOAL (1) / S4_ROOT_AUTO (2): root the trees at the goals / at the smaller endpoint set --
    # legitimate because the graph is undirected; identical loads wherever the shortest path is unique.
    # Default AUTO: same loads as the reference wherever shortest paths are unique, 30x fewer trees on the
    # BASELINE workloads (goals are a 2 % node sample).  Set 0 for the reference's exact grouping.
    "root_mode": 2,
    # Reproduce python-graph's per-orientation attribute lists: a street inserted as (u, v) with u > v
    # never shows a changed speed limit to the weight loop (streetnetwork.py:79-82 vs :113-125).
    "orientation_aliasing": True,
    # Draw the trips on the GPU (Philox sampler, seed = random_seed + 37 * rank) instead of with random.sample on the
    # host: same distribution (one uniform origin, one uniform goal per resident), a different but reproducible
    # list -- trip lists are run inputs -- and no Python loop over the residents.  For 10^7..10^8-trip runs.
    "sample_trips_on_device": False,
    # Warn (RuntimeWarning + log line) when trips could not be walked back to their root: only possible on a
    # zero-weight plateau of more than 32 nodes; the reference would have raised KeyError (simulation.py:85).
    "warn_stuck": True,
    # SSSP kernel tunables (None = library default), see include/s4mpi.h s4_ctx_set_option
    "delta_factor": None,
    "block_threads": None,
    "ctas_per_sm": None,
    "pool_bytes": None,
    "batch_roots": None,
}
