"""Wire-format writer for the reference's `weather.proto` slice stream.

The reference declares the messages (`src/proto/weather.proto:57-118`: AtmosphericCell, AtmosphericSlice,
WeatherSimUpdate) but never compiles or fills them. This module emits them without protoc: protobuf's wire
format is small enough to encode directly, and doing it with numpy keeps a 1024x1024 slice at a few tens of
milliseconds. Field mapping (the proto's names on the left):

    temperature <- T      pressure <- p      humidity <- q
    wind_velocity_x <- u  wind_velocity_y <- v          (wind_velocity_z, precipitation_rate, cloud_density: unset)

A stream is a sequence of length-delimited `WeatherSimUpdate` messages (varint size, then the message), the
usual framing for protobuf streams. Pure numpy: importable without the CUDA extension.
"""
import numpy as np

_WT_VARINT, _WT_I64, _WT_LEN = 0, 1, 2


def _varint(n):
    n = int(n)
    if n < 0:
        n += 1 << 64  # int32/int64 fields: negative values are sign-extended to 64 bits
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        out.append(b | (0x80 if n else 0))
        if not n:
            return bytes(out)


def _tag(field, wire_type):
    return _varint((field << 3) | wire_type)


def _len_delimited(field, payload):
    return _tag(field, _WT_LEN) + _varint(len(payload)) + payload


def _double(field, value):
    return _tag(field, _WT_I64) + np.float64(value).tobytes()


# one AtmosphericCell with fields 1..5 always present: 5 x (1 tag byte + 8 bytes) = 45 bytes, wrapped as element
# of `repeated AtmosphericCell cells = 2` -> tag 0x12, length 45. Fixed size, so a slice is one structured array.
_CELL_DTYPE = np.dtype([("tag", "u1"), ("len", "u1")] +
                       [(f"{n}{k}", t) for n in ("t", "p", "q", "u", "v") for k, t in (("_tag", "u1"), ("", "<f8"))])
_CELL_TAGS = {"t": 0x09, "p": 0x11, "q": 0x19, "u": 0x21, "v": 0x29}  # (field << 3) | 1 for fields 1..5
assert _CELL_DTYPE.itemsize == 47


def encode_slice(z_level, u, v, temperature, pressure, humidity):
    """AtmosphericSlice (weather.proto:69-74) of one level; the arrays are (H, W), any float dtype."""
    u = np.asarray(u)
    h, w = u.shape
    cells = np.empty(h * w, _CELL_DTYPE)
    cells["tag"], cells["len"] = 0x12, 45
    for key, arr in (("t", temperature), ("p", pressure), ("q", humidity), ("u", u), ("v", v)):
        a = np.asarray(arr)
        if a.shape != (h, w):
            raise ValueError("Array dimensions must match field dimensions")
        cells[key + "_tag"] = _CELL_TAGS[key]
        cells[key] = a.reshape(-1)
    return (_tag(1, _WT_VARINT) + _varint(z_level) + cells.tobytes() +
            _tag(3, _WT_VARINT) + _varint(w) + _tag(4, _WT_VARINT) + _varint(h))


def encode_update(run_id, current_time, percent_complete, slice_bytes):
    """WeatherSimUpdate (weather.proto:104-118) around an encoded slice."""
    rid = run_id.encode()
    return (_len_delimited(1, rid) + _double(2, current_time) + _double(3, percent_complete) +
            _len_delimited(4, slice_bytes))


def frame(message):
    """Length-delimited framing of one message of a stream."""
    return _varint(len(message)) + message


def read_frames(data):
    """Split a stream written with frame() back into messages (for tests and small tools)."""
    out, i, n = [], 0, len(data)
    while i < n:
        size, shift = 0, 0
        while True:
            b = data[i]
            i += 1
            size |= (b & 0x7F) << shift
            shift += 7
            if not b & 0x80:
                break
        out.append(bytes(data[i:i + size]))
        i += size
    return out
