"""Offline heat maps from the .s4mpi files a run leaves behind (reference visualization.py:52-224).

Consumer of the hot path's outputs only: it reads ``street_network_<k>.s4mpi`` / ``traffic_load_<k>.s4mpi``
from the working directory and writes ``traffic_load_<k>.png``.  Same constructor, modes and colour modes as
the reference; written for Python 3 / Pillow >= 10 (``textsize`` is gone, ``filter`` is lazy).

  modes         COMPONENTS | TRAFFIC_LOAD | MAX_SPEED | IDEAL_SPEED | ACTUAL_SPEED
  colour modes  HEATMAP (dark blue .. red) | MONOCHROME (black .. white)

The two speed modes evaluate ``calculate_driving_speed`` (simulation.py:129-138) for every street; that is
the K1 arithmetic, so it runs through the device kernel in one batch (and needs a GPU, like K1 itself).
"""
from __future__ import annotations

import re
from os import listdir

from .persistence import persist_read


class Visualization(object):
    MODES = ("COMPONENTS", "TRAFFIC_LOAD", "MAX_SPEED", "IDEAL_SPEED", "ACTUAL_SPEED")
    COLOR_MODES = ("HEATMAP", "MONOCHROME")

    def __init__(self, street_network_filename_pattern, traffic_load_filename_pattern, mode="TRAFFIC_LOAD",
                 color_mode="HEATMAP", verbose=True):
        if mode not in self.MODES or color_mode not in self.COLOR_MODES:
            raise ValueError("unknown mode %r / colour mode %r" % (mode, color_mode))
        self.max_resolution = (2000, 2000)
        self.zoom = 1
        self.coord2km = (111.32, 66.4)      # km per degree of latitude / longitude around Hamburg (visualization.py:62)
        self.bounds = None
        self.street_network = None
        self.node_coords = dict()
        self.mode = mode
        self.color_mode = color_mode
        self.verbose = verbose
        self._component_of = None
        self.street_network_filename_expression = re.compile(street_network_filename_pattern)
        self.traffic_load_filename_expression = re.compile(traffic_load_filename_pattern)
        self.log("Welcome to Streets4MPI visualization!")
        self.log("Current display mode:", mode, "with color mode", color_mode)

    def log(self, *a):
        if self.verbose:
            print(*a)

    # ------------------------------------------------------------------ driver (visualization.py:67-141)
    def visualize(self):
        from PIL import Image, ImageDraw
        all_files = listdir(".")
        network_files = [f for f in all_files if self.street_network_filename_expression.search(f)]
        load_files = [f for f in all_files if self.traffic_load_filename_expression.search(f)]
        max_load = 0
        for f in load_files:                                  # one scale for every day of the run
            load = persist_read(f, is_array=True)
            if len(load):
                max_load = max(max_load, max(load))
        written = []
        step = 0
        # the reference loops until every load file is drawn; stop as well once no file can match any more
        last = max([int(m.group()) for f in load_files for m in [re.search(r"\d+", f)] if m] or [0])
        while load_files and step < last:
            step += 1
            self.log("Step counter", step)
            name = "street_network_%d.s4mpi" % step
            if name in network_files:
                self.log("  Found street network data, reading...")
                self._load_network(persist_read(name))
            name = "traffic_load_%d.s4mpi" % step
            if name in load_files:
                if self.street_network is None:
                    raise RuntimeError("no street network has been read before " + name)
                self.log("  Found traffic load data, reading and drawing...")
                load = persist_read(name, is_array=True)
                image = Image.new("RGBA", self.max_resolution, (0, 0, 0, 255))
                draw = ImageDraw.Draw(image)
                values = self._street_values(load, max_load)
                for (street, index, _length, _max_speed) in self.street_network:
                    if self.mode == "COMPONENTS":
                        color = "hsl(%d,100%%,50%%)" % (int(137.5 * self._component_of[street[0]]) % 360)
                    else:
                        color = self.value_to_color(values[index])
                    draw.line([self.node_coords[street[0]], self.node_coords[street[1]]], fill=color, width=1)
                image = self.image_finalize(image, max_load)
                out = "traffic_load_%d.png" % step
                self.log("  Saving image to disk (%s) ..." % out)
                image.save(out)
                written.append(out)
                load_files.remove(name)
        self.log("Done!")
        return written

    def _load_network(self, network):
        self.street_network = network
        self.bounds = network.bounds
        if self.bounds is None:
            lons, lats = zip(*(network.node_coordinates(n) for n in network.get_nodes()))
            self.bounds = ((min(lats), max(lats)), (min(lons), max(lons)))
        span = max((self.bounds[0][1] - self.bounds[0][0]) * self.coord2km[0],
                   (self.bounds[1][1] - self.bounds[1][0]) * self.coord2km[1])
        self.zoom = self.max_resolution[0] / span if span > 0 else 1.0
        self.node_coords = dict()
        for node in network.get_nodes():
            lon, lat = network.node_coordinates(node)
            y = (lat - self.bounds[0][0]) * self.coord2km[0] * self.zoom
            x = (lon - self.bounds[1][0]) * self.coord2km[1] * self.zoom
            self.node_coords[node] = (x, self.max_resolution[1] - y)           # x = longitude, y = latitude
        if self.mode == "COMPONENTS":
            self._component_of = self.calculate_components(network)

    def _street_values(self, load, max_load):
        """Per street_index value in [0, 1] for the current mode (visualization.py:108-121)."""
        net = self.street_network
        rows = sorted(((idx, length, ms) for (_s, idx, length, ms) in net), key=lambda r: r[0])
        if self.mode == "TRAFFIC_LOAD":
            return [1.0 * load[idx] / max_load if max_load else 0.0 for (idx, _l, _m) in rows]
        if self.mode == "MAX_SPEED":
            return [min(140, 1.0 * ms / 140) for (_i, _l, ms) in rows]
        if self.mode in ("IDEAL_SPEED", "ACTUAL_SPEED"):
            from .simulation import calculate_driving_speeds
            trips = [0] * len(rows) if self.mode == "IDEAL_SPEED" else [load[idx] for (idx, _l, _m) in rows]
            speeds = calculate_driving_speeds([r[1] for r in rows], [r[2] for r in rows], trips)
            return [min(1.0, 1.0 * v / 140) for v in speeds.tolist()]
        return [0.0] * len(rows)

    # ------------------------------------------------------------------ helpers
    @staticmethod
    def calculate_components(network):
        """{node: component number} (the reference used python-graph's connected_components)."""
        parent = {n: n for n in network.get_nodes()}

        def find(x):
            while parent[x] != x:
                parent[x] = parent[parent[x]]
                x = parent[x]
            return x

        for (street, _i, _l, _m) in network:
            a, b = find(street[0]), find(street[1])
            if a != b:
                parent[max(a, b)] = min(a, b)
        numbering = {}
        out = {}
        for n in network.get_nodes():
            out[n] = numbering.setdefault(find(n), len(numbering) + 1)
        return out

    def value_to_color(self, value):
        value = min(1.0, max(0.0, value))
        if self.color_mode == "MONOCHROME":
            brightness = min(255, int(15 + 240 * value))
            return (brightness, brightness, brightness, 0)
        if value <= 0.2:                                       # almost black to blue
            return "hsl(260,100%%,%d%%)" % (5 + int(45 * 5 * value))
        return "hsl(%d,100%%,50%%)" % int(260 * (1 - (value - 0.2) / 0.8))     # blue to red

    def image_finalize(self, image, max_load):
        """Crop, add the legend bar with its two captions and the credit line (visualization.py:165-217)."""
        from PIL import Image, ImageDraw, ImageFont
        image = self.auto_crop(image)
        white, black = (255, 255, 255, 0), (0, 0, 0, 0)
        padding = self.max_resolution[0] // 40
        legend = Image.new("RGBA", image.size, (0, 0, 0, 255))
        font = ImageFont.load_default()
        draw = ImageDraw.Draw(legend)

        def text_size(text):
            box = draw.textbbox((0, 0), text, font=font)
            return box[2] - box[0], box[3] - box[1]

        bar_outer = self.max_resolution[0] // 50
        bar_inner = min(bar_outer - 4, int(bar_outer * 0.85))
        bar_inner -= (bar_outer - bar_inner) % 4
        bar_offset = max(2, (bar_outer - bar_inner) // 2)
        if self.mode != "COMPONENTS":
            height = legend.size[1]
            draw.rectangle([(0, 0), (bar_outer, height - 1)], fill=white)
            border = bar_offset // 2
            draw.rectangle([(border, border), (bar_outer - border, height - 1 - border)], fill=black)
            for y in range(bar_offset, height - bar_offset):
                value = 1.0 * (y - bar_offset) / max(1, height - 2 * bar_offset)
                draw.line([(bar_offset, y), (bar_offset + bar_inner, y)], fill=self.value_to_color(1.0 - value))
            captions = {
                "TRAFFIC_LOAD": (str(round(max_load, 1)) + " cars gone through", "0 cars gone through"),
                "MAX_SPEED": ("speed limit: 140 km/h or higher", "speed limit: 0 km/h"),
                "IDEAL_SPEED": ("ideal driving speed: 140 km/h or higher", "ideal driving speed: 0 km/h"),
                "ACTUAL_SPEED": ("actual driving speed: 140 km/h or higher", "actual driving speed: 0 km/h"),
            }[self.mode]
            draw.text((int(bar_outer * 1.3), 0), captions[0], font=font, fill=white)
            draw.text((int(bar_outer * 1.3), height - text_size(captions[1])[1] - 2), captions[1], font=font, fill=white)
            legend = self.auto_crop(legend)
        else:
            legend = legend.crop((0, 0, 1, legend.size[1]))
        credit = ("Generated by streets4mpi_b200 (Streets4MPI file formats) using data from the OpenStreetMap project. "
                  "Licensed under CC-BY-SA 2.0 (https://creativecommons.org/licenses/by-sa/2.0/).")
        credit_h = text_size(credit)[1]
        final = Image.new("RGB", (image.size[0] + legend.size[0] + 3 * padding,
                                  max(image.size[1], legend.size[1]) + 2 * padding + credit_h + 4), black)
        final.paste(image, (padding, padding))
        final.paste(legend, (image.size[0] + 2 * padding, padding))
        ImageDraw.Draw(final).text((2, max(image.size[1], legend.size[1]) + 2 * padding), credit, font=font, fill=white)
        return final

    @staticmethod
    def auto_crop(image):
        """Remove the black margin (visualization.py:219-224)."""
        from PIL import Image, ImageChops
        empty = Image.new("RGBA", image.size, (0, 0, 0, 255))
        bbox = ImageChops.difference(image.convert("RGB"), empty.convert("RGB")).getbbox()
        return image.crop(bbox) if bbox else image


if __name__ == "__main__":
    Visualization("^street_network_[0-9]+.s4mpi$", "^traffic_load_[0-9]+.s4mpi$", mode="TRAFFIC_LOAD").visualize()
