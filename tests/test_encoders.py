"""ROS-free message encoders (gm_encode_pointcloud2 / gm_encode_marker_array): the bytes are read back here with an
independent ROS 1 deserialiser written from the message definitions (sensor_msgs/PointCloud2, visualization_msgs/Marker,
std_msgs/Header, geometry_msgs/Pose|Vector3|Point, std_msgs/ColorRGBA) and compared with what the reference puts into
those messages (src/geometric_mapping.cpp:100-117, src/tunnel_processing.cpp:171-203, :228-252, :260-300).
Host-side formatting only: runs without a GPU."""
import struct

import numpy as np

from geometric_mapping_b200 import capi


class Rd:
    def __init__(self, b):
        self.b, self.o = b, 0

    def take(self, fmt):
        v = struct.unpack_from("<" + fmt, self.b, self.o)
        self.o += struct.calcsize("<" + fmt)
        return v if len(v) > 1 else v[0]

    def string(self):
        n = self.take("I")
        s = self.b[self.o:self.o + n].decode()
        self.o += n
        return s

    def header(self):
        return {"seq": self.take("I"), "stamp": self.take("II"), "frame_id": self.string()}


def read_pointcloud2(b):
    r = Rd(b)
    m = {"header": r.header(), "height": r.take("I"), "width": r.take("I"), "fields": []}
    for _ in range(r.take("I")):
        m["fields"].append((r.string(), r.take("I"), r.take("B"), r.take("I")))
    m["is_bigendian"], m["point_step"], m["row_step"] = r.take("B"), r.take("I"), r.take("I")
    n = r.take("I")
    m["data"] = b[r.o:r.o + n]
    r.o += n
    m["is_dense"] = r.take("B")
    assert r.o == len(b)
    return m


def read_marker(r):
    m = {"header": r.header(), "ns": r.string(), "id": r.take("i"), "type": r.take("i"), "action": r.take("i"),
         "pose": r.take("7d"), "scale": r.take("3d"), "rgba": r.take("4f"), "lifetime": r.take("ii"), "frame_locked": r.take("B")}
    m["points"] = [r.take("3d") for _ in range(r.take("I"))]
    m["colors"] = [r.take("4f") for _ in range(r.take("I"))]
    m["text"], m["mesh_resource"], m["mesh_use_embedded_materials"] = r.string(), r.string(), r.take("B")
    return m


def test_pointcloud2_is_what_pcl_toROSMsg_builds():
    g = np.random.default_rng(1)
    cloud = g.normal(size=(1000, 4)).astype(np.float32)
    cloud[:, 3] = 1.0
    b = capi.encode_pointcloud2(cloud, "/velodyne", seq=7, stamp_ns=12 * 10**9 + 345)
    m = read_pointcloud2(b)
    assert m["header"] == {"seq": 7, "stamp": (12, 345), "frame_id": "/velodyne"}
    assert (m["height"], m["width"]) == (1, 1000)                       # unorganised cloud
    assert m["fields"] == [("x", 0, 7, 1), ("y", 4, 7, 1), ("z", 8, 7, 1)]  # FLOAT32 = 7
    assert (m["is_bigendian"], m["point_step"], m["row_step"], m["is_dense"]) == (0, 16, 16000, 1)
    assert m["data"] == cloud.tobytes()
    assert len(capi.encode_pointcloud2(np.empty((0, 4), np.float32))) == capi._lib().gm_pointcloud2_size(0, b"/velodyne")


def test_marker_arrays_carry_the_reference_arrow_payloads():
    fr = capi.gm_frame()
    fr.vals[:] = [0.5, 2.0, 3.0]
    fr.vecs[:] = [1, 0, 0, 0, 1, 0, 0, 0, 1]
    arrows = capi.markers_eigen(fr)
    b = capi.encode_marker_array(arrows, "eigenBasis", stamp_ns=5)
    r = Rd(b)
    n = r.take("I")
    ms = [read_marker(r) for _ in range(n)]
    assert r.o == len(b) and n == 3
    lam = np.abs(np.array([0.5, 2.0, 3.0], np.float32)) / np.linalg.norm(np.array([0.5, 2.0, 3.0], np.float32))
    for i, m in enumerate(ms):
        assert m["header"]["frame_id"] == "/velodyne" and m["header"]["seq"] == 0 and m["ns"] == "eigenBasis" and m["id"] == i
        assert (m["type"], m["action"]) == (0, 0)                       # ARROW, ADD
        assert m["pose"] == (0.0,) * 7 and m["lifetime"] == (0, 0) and m["frame_locked"] == 0 and m["colors"] == [] and m["text"] == ""
        assert m["points"][0] == (0.0, 0.0, 0.0) and m["points"][1] == tuple(float(v) for v in np.eye(3)[:, i])
        assert np.allclose(m["scale"], [0.1 - 0.05 * lam[i], 0.3 - 0.15 * lam[i], 0.25 - 0.125 * lam[i]], atol=1e-6)   # :274-278
        # colours are pushed as (a,r,g,b) (:199-202): alpha 1, then the unit colour of axis i
        want = [1.0 if i == 0 else 0.0, 1.0 if i == 1 else 0.0, 1.0 if i == 2 else 0.0, 1.0]
        assert list(m["rgba"]) == want
    # normals arrows: start = centroid, end = the normal itself (quirk B.4), blue (quirk B.5), ns "normals"
    cen = np.array([[1, 2, 3, 1], [4, 5, 6, 1]], np.float32)
    nn = np.zeros((2, 8), np.float32)
    nn[:, :3] = [[0, 0, 1], [0, 1, 0]]
    b = capi.encode_marker_array(capi.markers_normals(cen, nn), "normals")
    r = Rd(b)
    ms = [read_marker(r) for _ in range(r.take("I"))]
    assert len(ms) == 2 and ms[1]["points"] == [(4.0, 5.0, 6.0), (0.0, 1.0, 0.0)] and ms[0]["ns"] == "normals"
    assert np.allclose(ms[0]["scale"], [0.025, 0.075, 0.0625]) and list(ms[0]["rgba"]) == [0.0, 0.0, 1.0, 1.0]
    assert len(b) == capi._lib().gm_marker_array_size(2, b"/velodyne", b"normals")


def test_primitive_store_appends_reopens_and_drops_a_torn_record(tmp_path):
    """gm_store_*: append-only file of per-scan GMC1 blobs; the index is rebuilt on open and a torn tail is ignored."""
    import os

    path = str(tmp_path / "scans.gms")
    blobs = [bytes([i]) * (100 + 37 * i) for i in range(5)] + [b""]
    with capi.PrimitiveStore(path) as st:
        for i, b in enumerate(blobs):
            st.append_blob(b, scan_id=1000 + i, stamp_ns=10**9 * i, pose34=None if i % 2 else np.arange(12, dtype=np.float32))
        assert len(st) == 6
    with capi.PrimitiveStore(path, create=False) as st:
        assert len(st) == 6
        for i, b in enumerate(blobs):
            info = st.info(i)
            assert st.read(i) == b and info["scan_id"] == 1000 + i and info["stamp_ns"] == 10**9 * i and info["bytes"] == len(b)
            want = np.eye(3, 4, dtype=np.float32) if i % 2 else np.arange(12, dtype=np.float32).reshape(3, 4)
            assert np.array_equal(info["pose"], want)
    size = os.path.getsize(path)
    with open(path, "r+b") as f:          # tear the last two records: truncate inside record 4
        f.truncate(size - 150)
    with capi.PrimitiveStore(path, create=False) as st:
        assert len(st) == 4 and st.read(3) == blobs[3]
        st.append_blob(b"after the crash", scan_id=7)
        assert len(st) == 5
    with capi.PrimitiveStore(path, create=False) as st:
        assert len(st) == 5 and st.read(4) == b"after the crash" and st.info(4)["scan_id"] == 7
    with open(path, "r+b") as f:          # a flipped byte inside a blob fails the CRC: that record and what follows are dropped
        f.seek(8 + 76 + 10)
        f.write(b"\xff")
    with capi.PrimitiveStore(path, create=False) as st:
        assert len(st) == 0
