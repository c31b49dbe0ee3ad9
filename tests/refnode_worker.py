"""Worker process that hosts the reference's three lidar nodes (oracle/_ref/libref_*.so: src/scanRegistration.cpp,
src/laserOdometry.cpp, src/laserMapping.cpp compiled UNMODIFIED against oracle/shim) and moves messages between them the way
the ROS graph does.  TEST INFRASTRUCTURE.  Protocol: length-prefixed pickles on stdin/stdout (see tests/refnode_py.py).
The nodes keep their state in process globals and their threads never exit, hence one worker process per pipeline instance;
the worker leaves with os._exit."""
import ctypes as C
import os
import pickle
import struct
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")


class Node:
    def __init__(self, so, params, keep):
        self.lib = L = C.CDLL(os.path.join(REF, so))     # RTLD_LOCAL: the three nodes define the same global names
        L.refnode_start.argtypes = [C.c_char_p, C.c_char_p]
        L.refnode_push_cloud.argtypes = [C.c_char_p, C.c_double, C.c_void_p, C.c_long, C.c_int]
        L.refnode_push_odom.argtypes = [C.c_char_p, C.c_double, C.c_void_p]
        L.refnode_wait.restype = C.c_long
        L.refnode_wait.argtypes = [C.c_char_p, C.c_long, C.c_int]
        L.refnode_take_cloud.restype = C.c_long
        L.refnode_take_cloud.argtypes = [C.c_char_p, C.c_void_p, C.c_long, C.c_void_p]
        L.refnode_take_odom.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p]
        L.refnode_take_lm_trace.restype = C.c_long
        L.refnode_take_lm_trace.argtypes = [C.c_void_p, C.c_long]
        L.refnode_peek.restype = C.c_long
        L.refnode_peek.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_long]
        self.count = {}
        L.refnode_start(";".join(f"{k}={v}" for k, v in params.items()).encode(), ";".join(keep).encode())

    def push_cloud(self, topic, stamp, pts, dense=True):
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 4)
        self.lib.refnode_push_cloud(topic.encode(), stamp, pts.ctypes.data, len(pts), int(dense))

    def push_odom(self, topic, stamp, qt7):
        qt7 = np.ascontiguousarray(qt7, np.float64)
        self.lib.refnode_push_odom(topic.encode(), stamp, qt7.ctypes.data)

    def wait_next(self, topic, timeout_ms=120000):
        want = self.count.get(topic, 0) + 1
        got = self.lib.refnode_wait(topic.encode(), want, timeout_ms)
        if got < want:
            raise TimeoutError(f"reference node did not publish on {topic}")
        self.count[topic] = want

    def take_cloud(self, topic):
        n = self.lib.refnode_take_cloud(topic.encode(), None, 0, None)
        if n < 0:
            return None
        out = np.empty((max(n, 1), 4), np.float32)
        st = C.c_double(0)
        self.lib.refnode_take_cloud(topic.encode(), out.ctypes.data, max(n, 1), C.byref(st))
        return out[:n]

    def take_odom(self, topic):
        qt = np.zeros(7)
        return qt if self.lib.refnode_take_odom(topic.encode(), qt.ctypes.data, None) == 0 else None

    def peek_scanreg(self, n):
        """cloudCurvature / cloudSortInd / cloudNeighborPicked / cloudLabel [0, n) of the scanRegistration node (scanRegistration.cpp:66-69)."""
        out = {}
        for what, name, dt in ((0, "curvature", np.float32), (1, "sort_ind", np.int32), (2, "picked", np.int32), (3, "label", np.int32)):
            a = np.zeros(max(n, 1), dt)
            self.lib.refnode_peek(what, a.ctypes.data, None, n)
            out[name] = a[:n]
        return out

    def peek_cloud(self, what):
        n = self.lib.refnode_peek(what, None, None, 0)
        pts = np.zeros((max(n, 1), 4), np.float32)
        aux = np.zeros(max(n, 1), np.int32)
        self.lib.refnode_peek(what, pts.ctypes.data, aux.ctypes.data, max(n, 1))
        return pts[:n], aux[:n]

    def peek_cen(self):
        aux = np.zeros(3, np.int32)
        self.lib.refnode_peek(4, None, aux.ctypes.data, 0)
        return aux

    def take_lm_traces(self):
        out = []
        while True:
            n = self.lib.refnode_take_lm_trace(None, 0)
            if n < 0:
                return out
            buf = np.zeros((max(n, 1), 10))
            self.lib.refnode_take_lm_trace(buf.ctypes.data, max(n, 1))
            out.append(buf[:n].copy())


FEATS = ["/velodyne_cloud_2", "/laser_cloud_sharp", "/laser_cloud_less_sharp", "/laser_cloud_flat", "/laser_cloud_less_flat"]
ODO_OUT = ["/laser_cloud_corner_last", "/laser_cloud_surf_last", "/velodyne_cloud_3"]


class Pipeline:
    def __init__(self, n_scans, min_range, line_res, plane_res, skip_frame, want_maps, lvo_atan=False):
        self.skip = skip_frame
        self.reg = Node("libref_scan_registration_lvoatan.so" if lvo_atan else "libref_scan_registration.so", dict(scan_line=n_scans, minimum_range=min_range), FEATS)
        self.odo = Node("libref_laser_odometry.so", dict(mapping_skip_frame=skip_frame), ODO_OUT + ["/laser_odom_to_init", "/lvo_shim/lm_trace"])
        keep = ["/aft_mapped_to_init", "/aft_mapped_to_init_high_frec", "/velodyne_cloud_registered", "/lvo_shim/lm_trace"]
        if want_maps:
            keep += ["/laser_cloud_surround", "/laser_cloud_map"]
        self.map = Node("libref_laser_mapping.so", dict(mapping_line_resolution=line_res, mapping_plane_resolution=plane_res), keep)
        self.frame = 0
        self.t = [0.0, 0.0, 0.0]

    def registration(self, pts, dense=True):
        stamp = 1.0 + 0.1 * self.frame
        t0 = time.perf_counter()
        self.reg.push_cloud("/velodyne_points", stamp, pts, dense)
        self.reg.wait_next(FEATS[-1])
        self.t[0] = time.perf_counter() - t0
        return [self.reg.take_cloud(t) for t in FEATS]

    def odometry(self, feats):
        """feats: full, sharp, less_sharp, flat, less_flat.  Returns (odom pose, published clouds or None, lm traces)."""
        stamp = 1.0 + 0.1 * self.frame
        t0 = time.perf_counter()
        for topic, pts in zip(FEATS, feats):
            self.odo.push_cloud(topic, stamp, pts)
        self.odo.wait_next("/laser_odom_to_init")
        # the publications of laserOdometry.cpp:643-663 come after the odometry message of the same frame
        pub = (self.frame % self.skip) == 0
        if pub:
            self.odo.wait_next(ODO_OUT[-1])
        self.t[1] = time.perf_counter() - t0
        pose = self.odo.take_odom("/laser_odom_to_init")
        clouds = [self.odo.take_cloud(t) for t in ODO_OUT] if pub else None
        return pose, clouds, self.odo.take_lm_traces()

    def mapping(self, odom, clouds):
        """Returns dict(high_freq pose, and when the frame is mapped: pose, registered, lm traces, optional map clouds)."""
        stamp = 1.0 + 0.1 * self.frame
        out = {}
        t0 = time.perf_counter()
        if clouds is not None:
            for topic, pts in zip(ODO_OUT, clouds):
                self.map.push_cloud(topic, stamp, pts)
        self.map.push_odom("/laser_odom_to_init", stamp, odom)
        self.map.wait_next("/aft_mapped_to_init_high_frec")
        out["high_freq"] = self.map.take_odom("/aft_mapped_to_init_high_frec")
        if clouds is not None:
            self.map.wait_next("/aft_mapped_to_init")
            out["pose"] = self.map.take_odom("/aft_mapped_to_init")
            out["registered"] = self.map.take_cloud("/velodyne_cloud_registered")
            out["lm"] = self.map.take_lm_traces()
            for k, t in (("surround", "/laser_cloud_surround"), ("whole_map", "/laser_cloud_map")):
                c = self.map.take_cloud(t)
                if c is not None:
                    out[k] = c
        self.t[2] = time.perf_counter() - t0
        return out

    def step(self, pts, dense=True, detail=False):
        feats = self.registration(pts, dense)
        pose, clouds, lm_odo = self.odometry(feats)
        m = self.mapping(pose, clouds)
        self.frame += 1
        r = dict(odom=pose, map=m.get("pose"), high_freq=m["high_freq"], times=list(self.t))
        if detail:
            r.update(self.reg.peek_scanreg(len(feats[0])))
            if m.get("pose") is not None:
                r["corner_from_map"], r["surf_from_map"] = self.map.peek_cloud(2)[0], self.map.peek_cloud(3)[0]
                r["map_corner"], r["map_corner_cube"] = self.map.peek_cloud(0)
                r["map_surf"], r["map_surf_cube"] = self.map.peek_cloud(1)
                r["cen"] = self.map.peek_cen()
            r.update(feats=feats, odo_clouds=clouds, lm_odo=lm_odo, lm_map=m.get("lm"), registered=m.get("registered"),
                     surround=m.get("surround"), whole_map=m.get("whole_map"))
        return r


def main():
    inp, out = sys.stdin.buffer, os.fdopen(os.dup(1), "wb")
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)          # the nodes print to stdout (std::cout in laserOdometry.cpp:264-266, :350)
    pipe = None

    def reply(obj):
        b = pickle.dumps(obj, protocol=4)
        out.write(struct.pack("<Q", len(b))); out.write(b); out.flush()

    while True:
        h = inp.read(8)
        if len(h) < 8:
            break
        cmd, args = pickle.loads(inp.read(struct.unpack("<Q", h)[0]))
        try:
            if cmd == "create":
                pipe = Pipeline(*args); reply(("ok", None))
            elif cmd == "step":
                reply(("ok", pipe.step(*args)))
            elif cmd == "registration":
                reply(("ok", pipe.registration(*args)))
            elif cmd == "quit":
                reply(("ok", None)); break
            else:
                reply(("err", f"unknown command {cmd}"))
        except Exception as e:  # noqa: BLE001
            reply(("err", repr(e)))
    os._exit(0)


if __name__ == "__main__":
    main()
