"""OpenStreetMap XML -> StreetNetwork (reference osmdata.py:37-231), one-off ingest in front of the hot path.

The reference parsed with imposm.parser (a C extension that is not installable here); this module reads
OSM XML with the standard library and feeds the same four record streams the reference's callbacks
received -- coords (id, lon, lat) of every node, tagged nodes, ways (id, tags, refs), relations (id, tags,
[(ref, type, role)]) -- into the same construction rules:

  * every consecutive ref pair of every way carrying a ``highway`` tag becomes a street, first one wins
    (osmdata.py:103-133); footways & co. get 1 km/h so that their weight stays finite (:83-86);
  * speed limit = highway-class table (:71-86), default 50 (:119), overridden by ``maxspeed``: plain digits,
    "<n> mph" taken as the bare number (no unit conversion, :125-127), "none" -> 140 (:128-129);
  * length = haversine with R = 6,367,000 m (:218-231);
  * node categories = all descendant nodes of relations / ways / nodes tagged landuse=residential|
    industrial|commercial, intersected with the street network's nodes (:137-168, :190-204).
"""
from __future__ import annotations

import xml.etree.ElementTree as ET
from math import asin, cos, radians, sin, sqrt

from .streetnetwork import StreetNetwork

MAX_SPEED_BY_HIGHWAY = {
    "motorway": 140, "trunk": 120, "primary": 100, "secondary": 80, "tertiary": 70, "road": 50, "minor": 50,
    "unclassified": 50, "residential": 30, "track": 30, "service": 20, "path": 10,
    "cycleway": 1, "bridleway": 1, "pedestrian": 1, "footway": 1,
}


def haversine_m(lat1, lon1, lat2, lon2):
    lat1, lon1, lat2, lon2 = map(radians, [lat1, lon1, lat2, lon2])
    dlon = lon2 - lon1
    dlat = lat2 - lat1
    a = sin(dlat / 2) ** 2 + cos(lat1) * cos(lat2) * sin(dlon / 2) ** 2
    return 6367000 * (2 * asin(sqrt(a)))


def speed_limit(tags):
    limit = MAX_SPEED_BY_HIGHWAY.get(tags["highway"], 50)
    tag = tags.get("maxspeed")
    if tag is not None:
        if tag.isdigit():
            limit = int(tag)
        elif tag.endswith("mph"):
            limit = int(tag.replace("mph", "").strip(" "))
        elif tag == "none":
            limit = 140
    return limit


class GraphBuilder(object):
    LATITUDE = 0
    LONGITUDE = 1

    def __init__(self, osmfile):
        self.street_network = StreetNetwork()
        self.coords = dict()                      # id -> (lat, lon)
        self.bounds = {"min_lat": 9999, "max_lat": -9999, "min_lon": 9999, "max_lon": -9999}
        self.all_osm_relations = dict()
        self.all_osm_ways = dict()
        self.all_osm_nodes = dict()               # tagged nodes only, like imposm's nodes callback
        self.residential_nodes = set()
        self.industrial_nodes = set()
        self.commercial_nodes = set()
        self.connected_residential_nodes = set()
        self.connected_industrial_nodes = set()
        self.connected_commercial_nodes = set()
        self.max_speed_map = dict(MAX_SPEED_BY_HIGHWAY)
        self._parse(osmfile)

    def _parse(self, osmfile):
        for _event, el in ET.iterparse(osmfile, events=("end",)):
            if el.tag == "node":
                osmid, lat, lon = int(el.get("id")), float(el.get("lat")), float(el.get("lon"))
                self.coords[osmid] = (lat, lon)
                b = self.bounds
                b["min_lat"], b["max_lat"] = min(b["min_lat"], lat), max(b["max_lat"], lat)
                b["min_lon"], b["max_lon"] = min(b["min_lon"], lon), max(b["max_lon"], lon)
                tags = {t.get("k"): t.get("v") for t in el.findall("tag")}
                if tags:
                    self.all_osm_nodes[osmid] = (osmid, tags, (lon, lat))
                el.clear()
            elif el.tag == "way":
                osmid = int(el.get("id"))
                tags = {t.get("k"): t.get("v") for t in el.findall("tag")}
                refs = [int(n.get("ref")) for n in el.findall("nd")]
                self.all_osm_ways[osmid] = (osmid, tags, refs)
                el.clear()
            elif el.tag == "relation":
                osmid = int(el.get("id"))
                tags = {t.get("k"): t.get("v") for t in el.findall("tag")}
                members = [(int(m.get("ref")), m.get("type"), m.get("role")) for m in el.findall("member")]
                self.all_osm_relations[osmid] = (osmid, tags, members)
                el.clear()

    def build_street_network(self):
        net = self.street_network
        b = self.bounds
        if 9999 not in b.values() and -9999 not in b.values():
            net.set_bounds(b["min_lat"], b["max_lat"], b["min_lon"], b["max_lon"])
        for _osmid, tags, refs in self.all_osm_ways.values():
            if "highway" not in tags or not refs:
                continue
            limit = speed_limit(tags)
            for ref in refs:
                if not net.has_node(ref):
                    lat, lon = self.coords[ref]                 # KeyError on a dangling ref, like the reference
                    net.add_node(ref, lon, lat)
            for a, b_ in zip(refs[:-1], refs[1:]):
                if not net.has_street((a, b_)):
                    net.add_street((a, b_), self.length_haversine(a, b_), limit)
        return net

    def find_node_categories(self):
        buckets = {"residential": self.residential_nodes, "industrial": self.industrial_nodes,
                   "commercial": self.commercial_nodes}
        for table in (self.all_osm_relations, self.all_osm_ways, self.all_osm_nodes):
            for osmid, tags, _payload in table.values():
                bucket = buckets.get(tags.get("landuse"))
                if bucket is not None:
                    bucket |= self.get_all_child_nodes(osmid)
        on_network = set(self.street_network.get_nodes())
        self.connected_residential_nodes = self.residential_nodes & on_network
        self.connected_industrial_nodes = self.industrial_nodes & on_network
        self.connected_commercial_nodes = self.commercial_nodes & on_network

    def get_all_child_nodes(self, osmid, _seen=None):
        """All descendant node ids of an OSM object; lookup order node, relation, way (osmdata.py:190-204)."""
        if osmid in self.all_osm_nodes:
            return {osmid}
        if osmid in self.all_osm_relations:
            seen = _seen or set()
            if osmid in seen:                       # the reference would recurse forever on a cyclic relation
                return set()
            seen = seen | {osmid}
            children = set()
            for ref, _type, _role in self.all_osm_relations[osmid][2]:
                children |= self.get_all_child_nodes(ref, seen)
            return children
        if osmid in self.all_osm_ways:
            return set(self.all_osm_ways[osmid][2])
        return set()

    def length_haversine(self, id1, id2):
        (lat1, lon1), (lat2, lon2) = self.coords[id1], self.coords[id2]
        return haversine_m(lat1, lon1, lat2, lon2)
