"""``from mpi4py import MPI`` — the subset of the API used by hwang595/ps_pytorch:

    COMM_WORLD.Get_rank/Get_size/isend/irecv/recv/send/bcast/Isend/Irecv/Iprobe/Barrier,
    Request.wait/Wait/test, Request.Waitany/waitany (static), Status (.source/.tag), DOUBLE/FLOAT/INT, ANY_SOURCE/ANY_TAG.

Lower-case calls pickle Python objects, upper-case calls move raw buffers — the same split as mpi4py.  Environment:
REFSHIM_RANK, REFSHIM_SIZE, REFSHIM_DIR (directory holding one listening socket ``r<rank>.sock`` per process).
"""
import os
import pickle
import socket
import struct
import threading
import time
from collections import deque

import numpy as np

ANY_SOURCE = -1
ANY_TAG = -1
DOUBLE = "d"
FLOAT = "f"
INT = "i"
_HDR = struct.Struct("<iiBQ")          # source, tag, kind (0 = pickled object, 1 = raw buffer), payload bytes
_BCAST_TAG = -7777


class Status(object):
    def __init__(self):
        self.source = ANY_SOURCE
        self.tag = ANY_TAG
        self.count = 0

    def Get_source(self):
        return self.source

    def Get_tag(self):
        return self.tag


class _World(object):
    def __init__(self):
        self.rank = int(os.environ.get("REFSHIM_RANK", "0"))
        self.size = int(os.environ.get("REFSHIM_SIZE", "1"))
        self.dir = os.environ.get("REFSHIM_DIR", "/tmp/refshim")
        self.cv = threading.Condition()
        self.box = {}                  # (source, tag) -> deque of (kind, payload)
        self.order = deque()           # arrival order of keys, consumed by Waitany
        self.out = {}                  # dest -> (socket, lock)
        self.out_lock = threading.Lock()
        self._listen = None
        if self.size > 1:
            os.makedirs(self.dir, exist_ok=True)
            path = os.path.join(self.dir, "r%d.sock" % self.rank)
            if os.path.exists(path):
                os.unlink(path)
            s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            s.bind(path)
            s.listen(self.size + 4)
            self._listen = s
            t = threading.Thread(target=self._accept_loop, daemon=True)
            t.start()

    # ---------------------------------------------------------------- receive side
    def _accept_loop(self):
        while True:
            try:
                c, _ = self._listen.accept()
            except OSError:
                return
            threading.Thread(target=self._reader, args=(c,), daemon=True).start()

    @staticmethod
    def _read_exact(c, n):
        buf = bytearray(n)
        view = memoryview(buf)
        got = 0
        while got < n:
            k = c.recv_into(view[got:], n - got)
            if k == 0:
                raise EOFError
            got += k
        return buf

    def _reader(self, c):
        try:
            while True:
                src, tag, kind, n = _HDR.unpack(bytes(self._read_exact(c, _HDR.size)))
                payload = self._read_exact(c, n) if n else bytearray()
                with self.cv:
                    self.box.setdefault((src, tag), deque()).append((kind, payload))
                    self.order.append((src, tag))
                    self.cv.notify_all()
        except (EOFError, OSError):
            return

    # ---------------------------------------------------------------- send side
    def _conn(self, dest):
        with self.out_lock:
            ent = self.out.get(dest)
            if ent is None:
                path = os.path.join(self.dir, "r%d.sock" % dest)
                deadline = time.time() + 120.0
                while True:
                    try:
                        s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
                        s.connect(path)
                        break
                    except OSError:
                        s.close()
                        if time.time() > deadline:
                            raise
                        time.sleep(0.02)
                s.setsockopt(socket.SOL_SOCKET, socket.SO_SNDBUF, 8 << 20)
                ent = self.out[dest] = (s, threading.Lock())
            return ent

    def send_frame(self, dest, tag, kind, payload):
        if dest == self.rank:
            with self.cv:
                self.box.setdefault((self.rank, tag), deque()).append((kind, bytearray(payload)))
                self.order.append((self.rank, tag))
                self.cv.notify_all()
            return
        s, lock = self._conn(dest)
        mv = memoryview(payload).cast("B")
        with lock:
            s.sendall(_HDR.pack(self.rank, tag, kind, mv.nbytes))
            if mv.nbytes:
                s.sendall(mv)

    # ---------------------------------------------------------------- matching
    def _match(self, source, tag, pop):
        """Caller holds cv.  Returns (key, (kind, payload)) or None."""
        if source != ANY_SOURCE and tag != ANY_TAG:
            q = self.box.get((source, tag))
            if q:
                return (source, tag), (q.popleft() if pop else q[0])
            return None
        for key, q in self.box.items():
            if q and (source == ANY_SOURCE or key[0] == source) and (tag == ANY_TAG or key[1] == tag) and key[1] != _BCAST_TAG:
                return key, (q.popleft() if pop else q[0])
        return None


_W = None


def _world():
    global _W
    if _W is None:
        _W = _World()
    return _W


def _as_buffer(spec):
    buf = spec[0] if isinstance(spec, (list, tuple)) else spec
    return buf


class Request(object):
    def __init__(self, source=None, tag=None, kind=0, buf=None, done=False, value=None):
        self.source, self.tag, self.kind, self.buf = source, tag, kind, buf
        self.done, self.value, self.null = done, value, False

    def _complete(self, key, item, status):
        kind, payload = item
        if self.kind == 0:
            self.value = pickle.loads(payload) if kind == 0 else bytes(payload)
        else:
            dst = memoryview(self.buf).cast("B") if not isinstance(self.buf, np.ndarray) else memoryview(self.buf.reshape(-1).view(np.uint8))
            n = min(dst.nbytes, len(payload))
            dst[:n] = memoryview(payload)[:n]
            self.value = None
        self.done = True
        if status is not None:
            status.source, status.tag, status.count = key[0], key[1], len(payload)

    def wait(self, status=None):
        if self.done:
            self.null = True
            return self.value
        w = _world()
        with w.cv:
            while True:
                m = w._match(self.source, self.tag, True)
                if m is not None:
                    self._complete(m[0], m[1], status)
                    self.null = True
                    return self.value
                w.cv.wait(1.0)

    Wait = wait

    def test(self, status=None):
        if self.done:
            return True, self.value
        w = _world()
        with w.cv:
            m = w._match(self.source, self.tag, True)
            if m is None:
                return False, None
            self._complete(m[0], m[1], status)
            return True, self.value

    Test = test

    # ---- MPI.Request.Waitany(requests=..., status=...) / waitany: complete ONE active receive, in arrival order
    @staticmethod
    def _waitany(requests, status):
        w = _world()
        index = getattr(requests, "_refshim_index", None) if not isinstance(requests, list) else None
        cache = Request._cache
        ent = cache.get(id(requests))
        if ent is None or ent[0] is not requests or ent[1] != len(requests):
            table = {}
            for i, r in enumerate(requests):
                if r is not None and not r.null and not r.done:
                    table.setdefault((r.source, r.tag), deque()).append(i)
            cache.clear()
            cache[id(requests)] = ent = (requests, len(requests), table)
        table = ent[2]
        for i, r in enumerate(requests):                 # sends (already complete) count as completed requests
            if r is not None and r.done and not r.null:
                r.null = True
                return i, r.value
        if not table:
            return -32766, None                          # MPI.UNDEFINED: nothing active
        with w.cv:
            while True:
                while w.order:
                    key = w.order.popleft()
                    q = table.get(key)
                    if q and w.box.get(key):
                        i = q.popleft()
                        if not q:
                            del table[key]
                        r = requests[i]
                        r._complete(key, w.box[key].popleft(), status)
                        r.null = True
                        return i, r.value
                # the arrival may have been queued before these requests were posted: fall back to a scan
                for key in list(table.keys()):
                    if w.box.get(key):
                        q = table[key]
                        i = q.popleft()
                        if not q:
                            del table[key]
                        r = requests[i]
                        r._complete(key, w.box[key].popleft(), status)
                        r.null = True
                        return i, r.value
                w.cv.wait(1.0)

    _cache = {}

    @staticmethod
    def Waitany(requests, status=None):
        return Request._waitany(requests, status)[0]

    @staticmethod
    def waitany(requests, status=None):
        return Request._waitany(requests, status)

    @staticmethod
    def Waitall(requests, statuses=None):
        for r in requests:
            r.wait()
        return True

    waitall = Waitall


class Intracomm(object):
    def Get_rank(self):
        return _world().rank

    def Get_size(self):
        return _world().size

    rank = property(Get_rank)
    size = property(Get_size)

    # ---- pickled objects
    def send(self, obj, dest, tag=0):
        _world().send_frame(dest, tag, 0, pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL))

    def isend(self, obj, dest, tag=0):
        self.send(obj, dest, tag)
        return Request(done=True)

    def irecv(self, buf=None, source=ANY_SOURCE, tag=ANY_TAG):
        return Request(source=source, tag=tag, kind=0)

    def recv(self, buf=None, source=ANY_SOURCE, tag=ANY_TAG, status=None):
        return Request(source=source, tag=tag, kind=0).wait(status)

    def bcast(self, obj, root=0):
        w = _world()
        if w.size == 1:
            return obj
        if w.rank == root:
            payload = pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)
            for d in range(w.size):
                if d != root:
                    w.send_frame(d, _BCAST_TAG, 0, payload)
            return obj
        return Request(source=root, tag=_BCAST_TAG, kind=0).wait()

    # ---- raw buffers
    def Send(self, spec, dest, tag=0):
        buf = _as_buffer(spec)
        arr = np.ascontiguousarray(buf) if isinstance(buf, np.ndarray) else buf
        _world().send_frame(dest, tag, 1, memoryview(arr).cast("B") if not isinstance(arr, np.ndarray) else arr.reshape(-1).view(np.uint8).data)

    def Isend(self, spec, dest, tag=0):
        self.Send(spec, dest, tag)
        return Request(done=True)

    def Irecv(self, spec, source=ANY_SOURCE, tag=ANY_TAG):
        return Request(source=source, tag=tag, kind=1, buf=_as_buffer(spec))

    def Recv(self, spec, source=ANY_SOURCE, tag=ANY_TAG, status=None):
        return Request(source=source, tag=tag, kind=1, buf=_as_buffer(spec)).wait(status)

    def Iprobe(self, source=ANY_SOURCE, tag=ANY_TAG, status=None):
        w = _world()
        with w.cv:
            m = w._match(source, tag, False)
        if m is None:
            return False
        if status is not None:
            status.source, status.tag = m[0]
        return True

    iprobe = Iprobe

    def Barrier(self):
        w = _world()
        if w.size == 1:
            return
        if w.rank == 0:
            for s in range(1, w.size):
                Request(source=s, tag=-7001, kind=0).wait()
            for s in range(1, w.size):
                self.send(None, s, -7002)
        else:
            self.send(None, 0, -7001)
            Request(source=0, tag=-7002, kind=0).wait()

    barrier = Barrier


COMM_WORLD = Intracomm()
UNDEFINED = -32766


def Wtime():
    return time.time()
