""".s4mpi files (reference persistence.py:28-59), Python 3.

  traffic_load_<day>.s4mpi    zlib(raw little-endian uint32[S]), no header     (is_array=True)
  street_network_<k>.s4mpi    zlib(pickle(StreetNetwork))

The reference opened the files in text mode under Python 2 (bytes == str there);
here they are binary.  Pickles use protocol 2 so a Python-2 reader can load them.
"""
import array
import pickle
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
