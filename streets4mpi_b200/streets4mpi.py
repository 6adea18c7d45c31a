"""Driver: the day loop and the rank-level exchange (reference streets4mpi.py:42-122).

``Streets4MPI()`` runs the whole program from its constructor, like the
reference.  Ranks are torch.distributed ranks (one per GPU, launched with
torchrun) instead of MPI ranks; without a launcher a single process runs the
same path serially (README.md:73-80).  The street network comes from
``settings["osm_file"]``: a ``synthetic:...`` spec (streets4mpi_b200.synth) or an
OpenStreetMap XML file through ``osmdata.GraphBuilder``.
"""
from __future__ import annotations

import random
from datetime import datetime

from . import distributed, synth
from .persistence import AsyncWriter
from .settings import device_settings, settings
from .simulation import Simulation
from .tripgenerator import TripGenerator


class Streets4MPI(object):

    def __init__(self):
        self.process_rank, number_of_processes = distributed.world()
        if number_of_processes > 1:
            # streets4mpi.py:44-46: join the job the launcher started (torchrun) and create the communicator of the
            # daily exchange; a rank that skipped this would silently route on its partial load
            from . import engine
            self.process_rank, number_of_processes = distributed.init(engine.default_device())

        self.log("Welcome to Streets4MPI!")
        random.seed(distributed.rank_seed(settings["random_seed"], self.process_rank))      # :50-51

        osm_file = settings["osm_file"]
        if str(osm_file).startswith("synthetic:"):
            self.log("Generating synthetic street network...")
            arrays = synth.parse_spec(osm_file)
            street_network = synth.network_from(arrays)
            cats = synth.node_categories(arrays["node_ids"], seed=settings["random_seed"])
            residential = cats["residential"].tolist()
            goals = cats["goals"].tolist()
        else:
            self.log("Reading OpenStreetMap data...")
            from .osmdata import GraphBuilder                                               # :54
            data = GraphBuilder(osm_file)
            self.log("Building street network...")
            street_network = data.build_street_network()                                    # :57
            self.log("Locating area types...")
            data.find_node_categories()                                                     # :64
            residential = data.connected_residential_nodes
            goals = data.connected_commercial_nodes | data.connected_industrial_nodes       # :74

        # same files as the reference, compressed and written by a worker thread while the next day runs
        writer = AsyncWriter() if self.process_rank == 0 and settings["persist_traffic_load"] else None
        if writer:
            self.log_indent("Saving street network to disk...")
            writer.write("street_network_1.s4mpi", street_network)                          # :59-61

        self.log("Generating trips...")
        number_of_residents = distributed.residents_of_rank(settings["number_of_residents"], number_of_processes)
        potential_origins = residential if settings["use_residential_origins"] else street_network.get_nodes()
        if device_settings.get("sample_trips_on_device"):
            trips = TripGenerator().sample_trips_on_device(number_of_residents, potential_origins, goals,
                                                           distributed.rank_seed(settings["random_seed"], self.process_rank))
        else:
            trips = TripGenerator().generate_trips(number_of_residents, potential_origins, goals)   # :75

        jam_tolerance = random.random()                                                     # :78
        self.log("Setting traffic jam tolerance to", str(round(jam_tolerance, 2)) + "...")

        simulation = Simulation(street_network, trips, jam_tolerance, self.log_indent)      # :82
        self.simulation = simulation

        for step in range(settings["max_simulation_steps"]):                                # :84
            if step > 0 and step % settings["steps_between_street_construction"] == 0:      # :86
                self.log_indent("Road construction taking place...")
                simulation.road_construction()
                if writer:
                    writer.write("street_network_" + str(step + 1) + ".s4mpi", simulation.street_network)

            self.log("Running simulation step", step + 1, "of", str(settings["max_simulation_steps"]) + "...")
            simulation.step()

            self.log("Exchanging traffic load data between nodes...")
            simulation.exchange()                                                           # :97-100, on the device

            if writer:
                self.log_indent("Saving traffic load to disk...")
                writer.write("traffic_load_" + str(step + 1) + ".s4mpi", simulation.traffic_load, is_array=True)

        if writer:
            writer.close()
        self.log("Done!")

    def log(self, *output):
        if settings["logging"] == "stdout":
            print("[ %s ][ p%d ]" % (datetime.now(), self.process_rank), *output)

    def log_indent(self, *output):
        if settings["logging"] == "stdout":
            print("[ %s ][ p%d ]  " % (datetime.now(), self.process_rank), *output)


if __name__ == "__main__":
    Streets4MPI()
