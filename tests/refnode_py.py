"""Client of tests/refnode_worker.py: one reference pipeline (the reference's own three node sources compiled unmodified into
oracle/_ref against shim ROS / PCL / Ceres headers) per worker process.  TEST INFRASTRUCTURE: used by tests/, tests/golden/
generators and bench.py's reference legs only.  Needs oracle/_ref/libref_*.so (built where /root/reference exists)."""
import os
import pickle
import struct
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFNODE_LIBS = [os.path.join(ROOT, "oracle", "_ref", n) for n in ("libref_scan_registration.so", "libref_laser_odometry.so", "libref_laser_mapping.so")]


def available():
    return all(os.path.exists(p) for p in REFNODE_LIBS)


class RefPipeline:
    def __init__(self, n_scans=64, min_range=5.0, line_res=0.4, plane_res=0.8, skip_frame=1, want_maps=False, lvo_atan=False):
        """lvo_atan=True: the scanRegistration node built with the deterministic atan / atan2 of csrc/lvo_math.h in place of glibc's (see
        oracle/refnode/refnode.cpp) — the build the restatement must match bit for bit."""
        if not available():
            raise FileNotFoundError("oracle/_ref/libref_*.so not built (make -C oracle ref, needs /root/reference)")
        self.p = subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "refnode_worker.py")], stdin=subprocess.PIPE, stdout=subprocess.PIPE)
        self._call("create", (n_scans, min_range, line_res, plane_res, skip_frame, want_maps, lvo_atan))

    def _call(self, cmd, args=()):
        b = pickle.dumps((cmd, args), protocol=4)
        self.p.stdin.write(struct.pack("<Q", len(b))); self.p.stdin.write(b); self.p.stdin.flush()
        h = self.p.stdout.read(8)
        if len(h) < 8:
            raise RuntimeError("reference worker died")
        st, val = pickle.loads(self.p.stdout.read(struct.unpack("<Q", h)[0]))
        if st != "ok":
            raise RuntimeError(val)
        return val

    def step(self, pts, dense=True, detail=False):
        """One sweep through scanRegistration -> laserOdometry -> laserMapping.  Returns dict(odom, map (None when the frame is
        not mapped: mapping_skip_frame), high_freq, times[3] seconds, and with detail=True the intermediate clouds / LM traces)."""
        return self._call("step", (pts, dense, detail))

    def registration(self, pts, dense=True):
        return self._call("registration", (pts, dense))

    def close(self):
        if self.p and self.p.poll() is None:
            try:
                self._call("quit")
            except Exception:  # noqa: BLE001
                pass
            self.p.wait(timeout=10)
        self.p = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass
