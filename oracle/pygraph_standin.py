"""Stand-in for the subset of python-graph that Streets4MPI calls.

TEST INFRASTRUCTURE (oracle).  python-graph (``python-graph-core``; the
reference pins no version -- README.md:47-48 only names the project) is not
vendored in /root/reference and not installable here, so its behaviour is
RESTATED FROM MEMORY of the 1.8.x line: **parity unpinned**.

Call sites this has to satisfy (reference file:line):
  streetnetwork.py:24-25   imports ``graph`` and ``shortest_path``
  streetnetwork.py:37,45,54,61,65,69,80,91,95,99,105,109,114,123
  visualization.py:30,144  (connected_components -- not needed on the hot path)

Semantics that matter to the hot path:
  * ``graph`` is UNDIRECTED: ``add_edge((u, v))`` registers both orientations;
    neighbour lists keep insertion order.
  * edge attributes are kept as one *separate list per orientation*
    (``add_edge_attribute`` rebuilds ``old + [attr]`` for each key), so mutating
    the list of one orientation in place is invisible from the other one
    (matters for streetnetwork.py:79-82 vs :113-125).
  * ``set_edge_weight`` writes both orientations.
  * ``shortest_path`` is Dijkstra with strict ``<`` relaxation; two recalled
    variants are provided (sorted-list ``insort`` queue of >=1.8.1, linear-scan
    of <=1.8.0).  Both return ``(previous, dist)`` with the source mapped to
    ``None`` / ``0`` and unreachable nodes absent.
"""
from bisect import insort


class AdditionError(RuntimeError):
    pass


class graph(object):
    DIRECTED = False

    def __init__(self):
        self.node_neighbors = {}
        self.node_attr = {}
        self.edge_properties = {}
        self.edge_attr = {}

    # -- nodes ------------------------------------------------------------
    def add_node(self, node, attrs=None):
        if node in self.node_neighbors:
            raise AdditionError("Node %s already in graph" % (node,))
        self.node_neighbors[node] = []
        self.node_attr[node] = attrs if attrs is not None else []

    def has_node(self, node):
        return node in self.node_neighbors

    def nodes(self):
        return list(self.node_neighbors.keys())

    def node_attributes(self, node):
        return self.node_attr[node]

    def neighbors(self, node):
        return self.node_neighbors[node]

    def __getitem__(self, node):
        return iter(self.node_neighbors[node])

    def __iter__(self):
        return iter(self.nodes())

    def __len__(self):
        return len(self.node_neighbors)

    # -- edges ------------------------------------------------------------
    def has_edge(self, edge):
        u, v = edge
        return (u, v) in self.edge_properties and (v, u) in self.edge_properties

    def edges(self):
        return list(self.edge_properties.keys())

    def add_edge(self, edge, wt=1, label="", attrs=None):
        u, v = edge
        if v in self.node_neighbors[u] or u in self.node_neighbors[v]:
            raise AdditionError("Edge (%s, %s) already in graph" % (u, v))
        self.node_neighbors[u].append(v)
        if u != v:
            self.node_neighbors[v].append(u)
        for attr in (attrs or []):
            self.add_edge_attribute((u, v), attr)
        self._set_props((u, v), label=label, weight=wt)
        if u != v:
            self._set_props((v, u), label=label, weight=wt)

    def _set_props(self, edge, **props):
        self.edge_properties.setdefault(edge, {}).update(props)

    def add_edge_attribute(self, edge, attr):
        # one fresh list per orientation -- never shared
        self.edge_attr[edge] = self.edge_attributes(edge) + [attr]
        if edge[0] != edge[1]:
            rev = (edge[1], edge[0])
            self.edge_attr[rev] = self.edge_attributes(rev) + [attr]

    def edge_attributes(self, edge):
        try:
            return self.edge_attr[edge]
        except KeyError:
            return []

    def set_edge_weight(self, edge, wt):
        self._set_props(edge, weight=wt)
        if edge[0] != edge[1]:
            self._set_props((edge[1], edge[0]), weight=wt)

    def edge_weight(self, edge):
        return self.edge_properties[edge]["weight"]


def shortest_path(g, source):
    """Recalled >=1.8.1 variant: sorted list of (dist, node), lazy deletion."""
    dist = {source: 0}
    previous = {source: None}
    queue = [(0, source)]
    finished = set()
    while queue:
        du, u = queue.pop(0)
        if u in finished:
            continue
        finished.add(u)
        for v in g[u]:
            if v in finished:
                continue
            alt = du + g.edge_weight((u, v))
            if v not in dist or alt < dist[v]:
                dist[v] = alt
                previous[v] = u
                insort(queue, (alt, v))
    return previous, dist


def shortest_path_scan(g, source):
    """Recalled <=1.8.0 variant: O(V) scan for the closest unfinished node."""
    dist = {source: 0}
    previous = {source: None}
    pending = g.nodes()
    while pending:
        u = pending[0]
        for cand in pending[1:]:
            if u not in dist or (cand in dist and dist[cand] < dist[u]):
                u = cand
        pending.remove(u)
        if u not in dist:
            continue            # only unreachable nodes are left
        for v in g[u]:
            if v in pending:
                alt = dist[u] + g.edge_weight((u, v))
                if v not in dist or alt < dist[v]:
                    dist[v] = alt
                    previous[v] = u
    return previous, dist
