""".s4mpi files (reference persistence.py:28-59), Python 3.

  traffic_load_<day>.s4mpi    zlib(raw little-endian uint32[S]), no header     (is_array=True)
  street_network_<k>.s4mpi    zlib(pickle(StreetNetwork))

The reference opened the files in text mode under Python 2 (bytes == str there);
here they are binary.  The traffic loads are plain arrays any reader can parse; the pickled StreetNetwork
(protocol 2) names this package's class, so only this package loads the street_network files.
"""
import array
import pickle
import queue
import threading
import zlib


def persist_serialize(data, compress=True):
    blob = pickle.dumps(data, protocol=2)
    return zlib.compress(blob) if compress else blob


def persist_deserialize(data, compressed=True):
    return pickle.loads(zlib.decompress(data) if compressed else data)


def persist_write(filename, data, compress=True, is_array=False):
    if is_array:
        if not isinstance(data, array.array):
            data = array.array("I", data)
        blob = zlib.compress(data.tobytes())
    else:
        blob = persist_serialize(data, compress)
    with open(filename, "wb") as f:
        f.write(blob)


def persist_read(filename, compressed=True, is_array=False):
    with open(filename, "rb") as f:
        blob = f.read()
    if is_array:
        result = array.array("I")
        result.frombytes(zlib.decompress(blob))
        return result
    return persist_deserialize(blob, compressed)


class AsyncWriter(object):
    """persist_write off the critical path of the day loop (streets4mpi.py:89-90,102-104 wrote synchronously; at one
    simulated day per second the zlib pass over 8-240 MB of counters would be a visible share of it).

    ``write()`` snapshots the data in the caller -- a copy of the array bytes, or the pickle of the object as it is
    now -- and a worker thread compresses (zlib releases the GIL) and writes the files in submission order, byte for
    byte what persist_write produces.  ``close()`` waits for the queue to drain and re-raises the first error."""

    def __init__(self, max_pending=4):
        self._q = queue.Queue(max_pending)
        self._error = None
        self._thread = threading.Thread(target=self._run, name="s4mpi-persist", daemon=True)
        self._thread.start()

    def _run(self):
        while True:
            item = self._q.get()
            try:
                if item is None:
                    return
                filename, blob, compress = item
                if self._error is None:
                    with open(filename, "wb") as f:
                        f.write(zlib.compress(blob) if compress else blob)
            except Exception as e:                      # reported by close() / the next write()
                self._error = e
            finally:
                self._q.task_done()

    def write(self, filename, data, compress=True, is_array=False):
        if self._error is not None:
            raise self._error
        if is_array:
            blob = data.tobytes() if isinstance(data, array.array) else array.array("I", data).tobytes()
            compress = True                             # persist_write always compresses arrays
        else:
            blob = pickle.dumps(data, protocol=2)
        self._q.put((filename, blob, compress))

    def flush(self):
        self._q.join()
        if self._error is not None:
            raise self._error

    def close(self):
        if self._thread is not None:
            self._q.put(None)
            self._thread.join()
            self._thread = None
        if self._error is not None:
            err, self._error = self._error, None
            raise err

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

