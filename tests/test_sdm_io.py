"""the reference's .sdm files and export images, product vs oracle (Map::write/read map.cpp:490-575,
export.cpp:46-96).  The reference has no fixtures for the format; the CPU tests pin the byte layout stated in its sources."""
import os

import numpy as np
import pytest

from iris_lama_b200 import sdm

O = 1321122 * 32


def _room_cells(segments):
    cells = set()
    for x1, y1, x2, y2 in segments:
        n = int(max(abs(x2 - x1), abs(y2 - y1)) / 0.05) + 1
        for k in range(n + 1):
            cells.add((int((x1 + (x2 - x1) * k / n) * 20 + O + 0.5), int((y1 + (y2 - y1) * k / n) * 20 + O + 0.5)))
    return np.array(sorted(cells), np.uint32)


def _same_files(a, b, n_params=None):
    fa, fb = sdm.read_sdm(a, n_params), sdm.read_sdm(b, n_params)
    assert fa["header"].tobytes() == fb["header"].tobytes()
    assert fa["params"] == fb["params"]
    assert set(fa["patches"]) == set(fb["patches"])        # the patch ORDER is that of an unordered_map: not part of the format
    for pid, (cells, mask) in fa["patches"].items():
        assert cells.tobytes() == fb["patches"][pid][0].tobytes(), pid
        assert mask.tobytes() == fb["patches"][pid][1].tobytes(), pid
    assert os.path.getsize(a) == os.path.getsize(b)
    return fa


# ---- CPU: the oracle's writer follows the byte layout of the reference sources -------------------------------------
def test_oracle_sdm_layout_and_round_trip(po, tmp_path):
    d = po.DDM(0.05, 32, 0.5)
    cells = np.array([(O + 5, O + 7), (O + 40, O - 3), (O - 100, O + 64)], np.uint32)
    d.add(cells)
    d.update()
    p = tmp_path / "d.sdm"
    assert po.map_write("ddm", d, p)
    raw = np.fromfile(p, np.uint8)
    assert raw[:4].tobytes() == b".sdm"                                   # MAGIC 0x6d64732e, map.h:72
    assert raw[4:6].tobytes() == b"\x03\x01"                              # IO_VERSION 0x0103, map.h:75
    f = sdm.read_sdm(p)
    h = f["header"]
    assert (h["cell_size"], h["patch_length"], h["is_3d"]) == (10, 32, 0) and h["resolution"] == np.float32(0.05)
    assert np.frombuffer(f["params"], "<u4")[0] == 100                    # max_sqdist_ = ceil(0.5 * 20)^2
    assert raw.size == 32 + 4 + len(f["patches"]) * (8 + 10240 + 128)
    for x, y in cells:                                                    # obstacle cells: sqdist 0, valid, known
        c, m = f["patches"][(int(x) >> 5) * sdm.UNIVERSAL_CONSTANT + (int(y) >> 5)]
        ci = (int(x) & 31) | ((int(y) & 31) << 5)
        assert c[ci]["sqdist"] == 0 and c[ci]["valid"] == 1 and (int(m[ci >> 6]) >> (ci & 63)) & 1
    d2 = po.DDM(0.05, 32, 0.25)
    assert po.map_read("ddm", d2, p) and d2.max_sqdist == 100           # readParameters adopts the file's value
    q = tmp_path / "d2.sdm"
    assert po.map_write("ddm", d2, q)
    _same_files(p, q)
    img = po.map_image("ddm", d)
    n, mn, mx = d.bounds()
    assert img.shape == (mx[1] - mn[1], mx[0] - mn[0])
    assert img[int(cells[0][1]) - mn[1], int(cells[0][0]) - mn[0]] == 0 and img.max() == 255 and (img == 127).any()


def test_golden_sdm_file(po, tmp_path):
    """tests/golden/ddm_small.sdm (written by make_golden.py): the oracle still writes the same bytes, the host mirror reads them"""
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ddm_small.sdm")
    d = po.DDM(0.05, 32, 0.5)
    cells = np.array([(O + 5, O + 7), (O + 6, O + 7), (O + 40, O - 3), (O - 20, O + 30)], np.uint32)
    d.add(cells); d.update(); d.remove(cells[1:2]); d.update()
    p = tmp_path / "now.sdm"
    assert po.map_write("ddm", d, p)
    f = _same_files(p, gold)
    assert len(f["patches"]) == 7 and os.path.getsize(gold) == 36 + 7 * (8 + 10240 + 128)
    c, m = f["patches"][(int(O + 5) >> 5) * sdm.UNIVERSAL_CONSTANT + (int(O + 7) >> 5)]
    assert c[5 | (7 << 5)]["valid"] == 1 and c[5 | (7 << 5)]["sqdist"] == 0           # still an obstacle
    assert c[6 | (7 << 5)]["sqdist"] == 1 and c[6 | (7 << 5)]["ox"] == -1             # removed: now 1 cell from its neighbour


def test_product_sdm_writer_reader_and_images_on_the_host(po, synth, tmp_path):
    """the product's .sdm / image code (csrc/sdm_io.cpp, pure host code) fed with the oracle's cell planes: its files must equal
    the oracle's byte for byte, its reader must return the planes, its images must equal the oracle's"""
    import ctypes as C
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    subprocess.check_call(["make", "-C", os.path.join(here, "emu"), "-s"])
    L = C.CDLL(os.path.join(here, "emu", "_build", "libsdm_hooks.so"))
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    ds = synth.make_dataset("room", 6, n_beams=180)
    o = po.Slam2D(po.SlamOptions.defaults(trans_thresh=0.05, rot_thresh=0.05))
    o.set_pose(*ds.truth[0])
    for t in range(6):
        o.update(ds.scans[t], ds.odom[t])
    # distance map
    n, mn, mx = o.dm_bounds(); w, h = int(mx[0] - mn[0]), int(mx[1] - mn[1])
    d = o.export_dm(mn[0], mn[1], w, h)
    a, b = tmp_path / "product.sdm", tmp_path / "oracle.sdm"
    assert L.sdmtest_write_distance(str(a).encode(), C.c_float(0.05), C.c_uint32(100), C.c_uint32(int(mn[0])), C.c_uint32(int(mn[1])), C.c_int(w), C.c_int(h),
                                    vp(d["sqdist"]), vp(d["valid"]), vp(d["known"]), vp(d["ox"]), vp(d["oy"]), vp(d["queued"])) == 1
    hd = po.map_handle("slam_dm", o)
    assert po.map_write("ddm", hd, b)
    _same_files(a, b)
    win = np.zeros(4, np.uint32); msq = np.zeros(1, np.uint32)
    assert L.sdmtest_read_distance(str(b).encode(), vp(win), vp(msq), None, None, None, None, None, None) == 1
    assert win.tolist() == [int(mn[0]), int(mn[1]), w, h] and msq[0] == 100
    r = {k: np.zeros_like(v) for k, v in d.items()}
    assert L.sdmtest_read_distance(str(b).encode(), vp(win), vp(msq), vp(r["sqdist"]), vp(r["valid"]), vp(r["known"]), vp(r["ox"]), vp(r["oy"]), vp(r["queued"])) == 1
    for k in d:
        assert (r[k] == d[k]).all(), k
    img = np.zeros((h, w), np.uint8)
    L.sdmtest_distance_image(C.c_uint32(int(mn[0])), C.c_uint32(int(mn[1])), C.c_int(w), C.c_int(h), vp(d["sqdist"]), vp(d["valid"]), vp(d["known"]), C.c_uint32(100),
                             C.c_double(0.05), vp(img))
    assert (img == po.map_image("ddm", hd)).all()
    # frequency occupancy map
    n, mn, mx = o.occ_bounds(); w, h = int(mx[0] - mn[0]), int(mx[1] - mn[1])
    e = o.export_occ(mn[0], mn[1], w, h)
    assert L.sdmtest_write_frequency(str(a).encode(), C.c_float(0.05), C.c_uint32(int(mn[0])), C.c_uint32(int(mn[1])), C.c_int(w), C.c_int(h), vp(e["occupied"]),
                                     vp(e["visited"]), vp(e["known"])) == 1
    ho = po.map_handle("slam_occ", o)
    assert po.map_write("freq", ho, b)
    _same_files(a, b)
    img = np.zeros((h, w), np.uint8)
    L.sdmtest_frequency_image(C.c_uint32(int(mn[0])), C.c_uint32(int(mn[1])), C.c_int(w), C.c_int(h), vp(e["occupied"]), vp(e["visited"]), vp(e["known"]), vp(img))
    assert (img == po.map_image("freq", ho)).all()


def test_corrupt_sdm_files_are_refused_not_thrown(tmp_path):
    """Map::read returns false on a short file (map.cpp:565-568); the product's reader must do the same for a truncated patch list and
    for a header whose 64-bit num_patches promises more than the file holds (no allocation sized from it, no exception)"""
    import ctypes as C
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    subprocess.check_call(["make", "-C", os.path.join(here, "emu"), "-s"])
    L = C.CDLL(os.path.join(here, "emu", "_build", "libsdm_hooks.so"))
    gold = os.path.join(here, "golden", "ddm_small.sdm")
    raw = open(gold, "rb").read()
    win, msq = (C.c_uint32 * 4)(), C.c_uint32()

    def reads(data):
        p = tmp_path / "x.sdm"
        p.write_bytes(data)
        return L.sdmtest_read_distance(str(p).encode(), win, C.byref(msq), None, None, None, None, None, None)

    assert reads(raw) == 1
    assert reads(raw[:-100]) == 0                                   # last patch cut short
    assert reads(raw[:36]) == 0                                     # header + parameters only, 7 patches announced
    hdr = sdm.read_sdm(gold)["header"]
    off = hdr.dtype.fields["num_patches"][1]
    huge = bytearray(raw)
    huge[off:off + 8] = (2 ** 62).to_bytes(8, "little")             # would be a 4 EiB resize
    assert reads(bytes(huge)) == 0


def test_png_writer_round_trip(tmp_path):
    import struct
    import zlib
    from iris_lama_b200 import api
    g = (np.arange(35 * 50) % 251).astype(np.uint8).reshape(35, 50)
    api.write_png(tmp_path / "g.png", g)
    raw = open(tmp_path / "g.png", "rb").read()
    assert raw[:8] == b"\x89PNG\r\n\x1a\n" and struct.unpack(">II", raw[16:24]) == (50, 35)
    i = raw.index(b"IDAT")
    n = struct.unpack(">I", raw[i - 4:i])[0]
    rows = np.frombuffer(zlib.decompress(raw[i + 4:i + 4 + n]), np.uint8).reshape(35, 51)
    assert (rows[:, 0] == 0).all() and (rows[:, 1:] == g).all()


# ---- GPU: files and images of device maps equal the oracle's -------------------------------------------------------
@pytest.mark.gpu
def test_pf_maps_written_as_reference_sdm_files(gpu_api, po, synth, tmp_path):
    ds = synth.make_dataset("room", 8, n_beams=360)
    kw = dict(trans_thresh=0.05, rot_thresh=0.05, seed=11)
    g = gpu_api.PFSlam2D(gpu_api.PFSlam2D.Options(6, **kw))
    o = po.PFSlam2D(po.PFOptions.defaults(6, **kw))
    g.setPrior(*ds.truth[0]); o.set_prior(*ds.truth[0])
    for t in range(8):
        assert g.update(ds.scans[t], ds.odom[t]) == o.update(ds.scans[t], ds.odom[t])
    for particle in (0, g.getBestParticleIdx()):
        for kind, name, okind in ((0, "occ", "freq"), (1, "dm", "ddm")):
            a, b = tmp_path / ("g_%s.sdm" % name), tmp_path / ("o_%s.sdm" % name)
            g.writeMap(particle, kind, a)
            h = po.map_handle("pf_" + name, o, particle)
            assert po.map_write(okind, h, b)
            f = _same_files(a, b)
            assert len(f["patches"]) > 4
            assert (g.exportImage(particle, kind) == po.map_image(okind, h)).all()
    g.saveOccImage(tmp_path / "best.png")
    assert os.path.getsize(tmp_path / "best.png") > 100


@pytest.mark.gpu
@pytest.mark.parametrize("occupancy", [0, 1])
def test_slam2d_maps_written_as_reference_sdm_files(gpu_api, po, synth, tmp_path, occupancy):
    ds = synth.make_dataset("corridor", 8, n_beams=360)
    kw = dict(trans_thresh=0.05, rot_thresh=0.05)
    g = gpu_api.Slam2D(gpu_api.Slam2D.Options(occupancy=occupancy, **kw))
    o = (po.Slam2DProb if occupancy else po.Slam2D)(po.SlamOptions.defaults(**kw))
    g.setPose(*ds.truth[0]); o.set_pose(*ds.truth[0])
    for t in range(8):
        assert g.update(ds.scans[t], ds.odom[t]) == o.update(ds.scans[t], ds.odom[t])
    pre = "slamp_" if occupancy else "slam_"
    for kind, name, okind in ((0, "occ", "prob" if occupancy else "freq"), (1, "dm", "ddm")):
        a, b = tmp_path / ("g_%s.sdm" % name), tmp_path / ("o_%s.sdm" % name)
        g.writeMap(kind, a)
        h = po.map_handle(pre + name, o)
        assert po.map_write(okind, h, b)
        _same_files(a, b)
        assert (g.exportImage(kind) == po.map_image(okind, h)).all()


@pytest.mark.gpu
def test_distance_map_file_round_trip_through_the_device(gpu_api, po, tmp_path):
    rng = np.random.default_rng(5)
    cells = np.unique((rng.integers(-60, 60, (150, 2)) + O).astype(np.uint32), axis=0)
    od = po.DDM(0.05, 32, 1.0)
    od.add(cells); od.update(); od.remove(cells[:40]); od.update()
    a = tmp_path / "oracle.sdm"
    assert po.map_write("ddm", od, a)
    gd = gpu_api.DynamicDistanceMap(0.05, 32, 1.0)
    gd.read(a)                                            # Map::read into the device map
    b = tmp_path / "device.sdm"
    gd.write(b)
    _same_files(a, b)
    assert (gd.exportImage() == po.map_image("ddm", od)).all()
    more = (rng.integers(-60, 60, (30, 2)) + O).astype(np.uint32)     # the loaded map keeps working: same brushfire afterwards
    gd.addObstacle(more); od.add(more)
    assert gd.update() == od.update()
    gd.write(b); po.map_write("ddm", od, a)
    _same_files(a, b)
    with pytest.raises(gpu_api.LamaError):
        gpu_api.DynamicDistanceMap(0.05, 32, 0.5).read(a)  # another l2_max


@pytest.mark.gpu
def test_loc2d_occupancy_map_loaded_from_file(gpu_api, po, synth, tmp_path):
    ds = synth.make_dataset("loc_room", 3)
    kw = dict(trans_thresh=0.01, rot_thresh=0.01, gloc_particles=500)
    gl = gpu_api.Loc2D(gpu_api.Loc2D.Options(**kw))
    ol = po.Loc2D(po.LocOptions.defaults(**kw))
    xs = np.arange(int(-9.9 * 20), int(9.9 * 20))
    free = np.array([(x + O, y + O) for x in xs[::3] for y in xs[::3]], np.uint32)
    ol.occ_set(free, -1)
    p = tmp_path / "occ.sdm"
    assert po.map_write("simple", po.map_handle("loc_occ", ol), p)
    gl.occupancyRead(p)
    cells = _room_cells(ds.segments)
    gl.distance_map.addObstacle(cells); gl.distance_map.update()
    od = ol.dm(); od.add(cells); od.update()
    gl.setSeed(5); ol.set_seed(5)
    gl.setPose(0, 0, 0); ol.set_pose(0, 0, 0)
    gl.triggerGlobalLocalization(); ol.trigger_global_localization()
    for t in range(3):   # the rejection sampling reads the loaded map: same candidates, same winner
        assert gl.update(ds.scans[t], ds.odom[t], force_update=(t == 0)) == ol.update(ds.scans[t], ds.odom[t], force=(t == 0))
        assert np.abs(gl.state() - ol.get()[0]).max() < 1e-9
