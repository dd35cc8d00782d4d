"""CPU test: the hand-encoded `weather.proto` slice stream (weather_sim/proto_stream.py) parses
with google.protobuf against the reference's schema (src/proto/weather.proto:57-74,104-118, restated here as a
descriptor because the reference never compiles its .proto and protoc is not in the image)."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location(
    "proto_stream", os.path.join(ROOT, "nvidia-jetson-workload_b200", "weather_sim", "proto_stream.py"))
ps = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ps)


def schema():
    pb = pytest.importorskip("google.protobuf")  # noqa: F841
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    F = descriptor_pb2.FieldDescriptorProto
    fd = descriptor_pb2.FileDescriptorProto(name="weather_slice.proto", package="nvidia.jetson.workload.weather",
                                            syntax="proto3")
    cell = fd.message_type.add(name="AtmosphericCell")
    for i, n in enumerate(("temperature", "pressure", "humidity", "wind_velocity_x", "wind_velocity_y",
                           "wind_velocity_z", "precipitation_rate", "cloud_density"), 1):
        cell.field.add(name=n, number=i, type=F.TYPE_DOUBLE, label=F.LABEL_OPTIONAL)
    sl = fd.message_type.add(name="AtmosphericSlice")
    sl.field.add(name="z_level", number=1, type=F.TYPE_INT32, label=F.LABEL_OPTIONAL)
    sl.field.add(name="cells", number=2, type=F.TYPE_MESSAGE, label=F.LABEL_REPEATED,
                 type_name=".nvidia.jetson.workload.weather.AtmosphericCell")
    sl.field.add(name="width", number=3, type=F.TYPE_INT32, label=F.LABEL_OPTIONAL)
    sl.field.add(name="height", number=4, type=F.TYPE_INT32, label=F.LABEL_OPTIONAL)
    up = fd.message_type.add(name="WeatherSimUpdate")
    up.field.add(name="run_id", number=1, type=F.TYPE_STRING, label=F.LABEL_OPTIONAL)
    up.field.add(name="current_time", number=2, type=F.TYPE_DOUBLE, label=F.LABEL_OPTIONAL)
    up.field.add(name="percent_complete", number=3, type=F.TYPE_DOUBLE, label=F.LABEL_OPTIONAL)
    up.field.add(name="current_slice", number=4, type=F.TYPE_MESSAGE, label=F.LABEL_OPTIONAL,
                 type_name=".nvidia.jetson.workload.weather.AtmosphericSlice")
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = getattr(message_factory, "GetMessageClass", None)
    d = pool.FindMessageTypeByName("nvidia.jetson.workload.weather.WeatherSimUpdate")
    return get(d) if get else message_factory.MessageFactory(pool).GetPrototype(d)


def test_stream_parses_with_protobuf_and_round_trips():
    Update = schema()
    rng = np.random.default_rng(5)
    H, W = 7, 11
    frames = b""
    want = []
    for step in range(3):
        f = {k: rng.standard_normal((H, W)).astype(np.float32) for k in "uvtpq"}
        f["u"][0, 0] = 0.0  # zeros are written explicitly (fixed-size cells) and must parse as zeros
        sl = ps.encode_slice(2, f["u"], f["v"], f["t"], f["p"], f["q"])
        frames += ps.frame(ps.encode_update("run-7", 0.5 * step, 10.0 * step, sl))
        want.append(f)
    msgs = ps.read_frames(frames)
    assert len(msgs) == 3
    for step, (raw, f) in enumerate(zip(msgs, want)):
        m = Update()
        m.ParseFromString(raw)
        assert m.run_id == "run-7" and m.current_time == 0.5 * step and m.percent_complete == 10.0 * step
        s = m.current_slice
        assert (s.z_level, s.width, s.height, len(s.cells)) == (2, W, H, H * W)
        for name, key in (("temperature", "t"), ("pressure", "p"), ("humidity", "q"), ("wind_velocity_x", "u"),
                          ("wind_velocity_y", "v")):
            got = np.array([getattr(c, name) for c in s.cells]).reshape(H, W)
            assert np.array_equal(got, f[key].astype(np.float64)), name
        assert all(c.wind_velocity_z == 0.0 and c.cloud_density == 0.0 for c in s.cells)


def test_shape_mismatch_is_an_error():
    a = np.zeros((3, 4), np.float32)
    with pytest.raises(ValueError):
        ps.encode_slice(0, a, a, a, np.zeros((4, 3), np.float32), a)
