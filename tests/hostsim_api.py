"""ctypes wrapper over tests/hostsim/libhostsim.so: the device interpreter's source compiled for the
CPU (unit-test harness only; the product has no CPU path)."""
import ctypes as C
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(os.path.join(ROOT, "tests", "hostsim", "libhostsim.so"))
        L.hs_new.restype = C.c_void_p
        L.hs_trace_count.restype = C.c_uint32
        _lib = L
    return _lib


def to_json(models, connectors):
    return (json.dumps([m.to_dict() for m in models]), json.dumps([c.to_dict() for c in connectors]))


class HostSim:
    def __init__(self, models, connectors, n=1, rep_offset=0, seed=42, rng_kind=0, instrument=True, stat_model_id=None,
                 queue_capacity=0, max_pending=0, trace_reps=0, trace_cap=0, raw_json=None, batch_warmup=0, batch_size=0):
        self.L = lib()
        mj, cj = raw_json if raw_json else to_json(models, connectors)
        err = C.create_string_buffer(512)
        status = C.c_int()
        h = self.L.hs_new(mj.encode(), cj.encode(), C.c_uint64(n), C.c_uint64(rep_offset), C.c_uint64(seed), rng_kind,
                          int(instrument), stat_model_id.encode() if stat_model_id else None, queue_capacity,
                          max_pending, trace_reps, trace_cap, batch_warmup, batch_size, err, 512, C.byref(status))
        self.status = status.value
        self.error = err.value.decode()
        self.h = C.c_void_p(h) if h else None
        self.n = n
        self.n_models = len(models) if models is not None else 0
        self.n_connectors = len(connectors) if connectors is not None else 0
        if self.h is None:
            raise RuntimeError("lowering failed (%d): %s" % (self.status, self.error))

    def __del__(self):
        try:
            if self.h:
                self.L.hs_free(self.h)
        except Exception:
            pass

    def set_reload_every(self, n):
        """store + reload the lane state every n steps, like consecutive kernel launches of K = n"""
        self.L.hs_set_reload_every(self.h, C.c_uint64(n))

    def set_rng_slots(self, n):
        """Philox draws cached per replication by the baked-shape code path (8 fused, 2 streaming launches)"""
        self.L.hs_set_rng_slots(self.h, C.c_uint32(n))

    def use_shape(self, k):
        """run through the literal-index (baked shape) code path; False if the network does not match shape k"""
        return bool(self.L.hs_use_shape(self.h, k))

    def set_stream(self, draws):  # draws[i, r]
        a = np.ascontiguousarray(draws, dtype=np.uint64)
        self.L.hs_set_stream(self.h, a.ctypes.data_as(C.POINTER(C.c_uint64)), C.c_uint64(a.shape[0]))

    def step_n(self, n):
        self.L.hs_step_n(self.h, C.c_uint64(n))

    def step_until(self, t):
        self.L.hs_step_until(self.h, C.c_double(t))

    def inject_input(self, target_id, target_port, content):
        return self.L.hs_inject(self.h, target_id.encode(), target_port.encode(), content.encode())

    def time(self):
        out = np.zeros(self.n, np.float64)
        self.L.hs_get_time(self.h, out.ctypes.data_as(C.POINTER(C.c_double)))
        return out

    def until(self, m):
        out = np.zeros(self.n, np.float64)
        self.L.hs_get_until(self.h, m, out.ctypes.data_as(C.POINTER(C.c_double)))
        return out

    def _row(self, which):
        out = np.zeros(self.n, np.uint32)
        self.L.hs_get_row(self.h, which, out.ctypes.data_as(C.POINTER(C.c_uint32)))
        return out

    def steps(self):
        return self._row(0)

    def events(self):
        return self._row(1)

    def draws(self):
        return self._row(2)

    def errors(self):
        return self._row(3)

    def npend(self):
        return self._row(4)

    def hash(self):
        out = np.zeros(self.n, np.uint64)
        self.L.hs_get_hash(self.h, out.ctypes.data_as(C.POINTER(C.c_uint64)))
        return out

    def wait(self):
        s = np.zeros(self.n, np.float64)
        n = np.zeros(self.n, np.uint32)
        self.L.hs_get_wait(self.h, s.ctypes.data_as(C.POINTER(C.c_double)), n.ctypes.data_as(C.POINTER(C.c_uint32)))
        return s, n

    def batch(self):
        """(sum of batch means, sum of squared batch means, running sum of the open batch) per replication"""
        a, b, c = (np.zeros(self.n, np.float64) for _ in range(3))
        p = lambda x: x.ctypes.data_as(C.POINTER(C.c_double))
        self.L.hs_get_batch(self.h, p(a), p(b), p(c))
        return a, b, c

    def conn_counts(self):
        out = np.zeros(max(1, self.n_connectors), np.uint64)
        self.L.hs_get_conn_counts(self.h, out.ctypes.data_as(C.POINTER(C.c_uint64)))
        return out[:self.n_connectors]

    def record_counts(self):
        out = np.zeros(self.n_models * 14, np.uint64)
        self.L.hs_get_record_counts(self.h, out.ctypes.data_as(C.POINTER(C.c_uint64)))
        return out.reshape(self.n_models, 14)

    def render(self, code):
        buf = C.create_string_buffer(256)
        self.L.hs_render(self.h, C.c_uint32(code), buf, 256)
        return buf.value.decode()

    def trace(self, rep=0):
        n = int(self.L.hs_trace_count(self.h, C.c_uint64(rep)))
        f = (C.c_uint32 * 6)()
        t = C.c_double()
        out = []
        for i in range(n):
            self.L.hs_trace_get(self.h, C.c_uint64(rep), i, f, C.byref(t))
            out.append(dict(step=f[0], kind=f[1], index=f[2], action=f[3], aux=f[4], content=self.render(f[5]),
                            time=t.value))
        return out

    def layout(self):
        out = (C.c_uint32 * 8)()
        self.L.hs_layout(self.h, out)
        return dict(n_drows=out[0], n_wrows=out[1], max_pending=out[2], ring_w_slots=out[3], ring_d_slots=out[4],
                    n_routes=out[5], n_models=out[6], n_connectors=out[7])

    def pending(self, rep=0):
        n = int(self.npend()[rep])
        o = (C.c_uint32 * 2)()
        out = []
        for i in range(n):
            self.L.hs_pending_get(self.h, C.c_uint64(rep), i, o)
            out.append((self.render(o[0]), o[1] & 0xFF, (o[1] >> 8) & 0xFF, (o[1] >> 16) & 0xFF, o[1] >> 24))
        return sorted(out, key=lambda p: p[4])  # emission order (the mailbox itself is sorted by target model)
