"""CPU test of the stage class's stream / event / buffer-rotation schedule (gs2mesh_b200/renderer.py::render_image_pair).

The REAL `render_image_pair` runs here against fake CUDA streams and events that record what was enqueued where and which
waits were issued.  From that record a happens-before graph is built (stream order, event waits, host synchronisations) and
the schedule is checked for data-race freedom the way a race detector would: every two accesses to the same buffer, at
least one of them a write, must be ordered by the graph.  Covered: 1-3 pairs in flight, 0-2 spare buffer sets, frame
copies on the render streams or on their own streams, and the consumers the product has --
  * the device-resident loop (bench.py / TSDF in memory): the caller's stream reads depth / final_T / left_u8 right after
    each call;
  * the host-buffer loops of bench.py (both orders of BENCH_E2E_ORDER) and tests/test_zz_gpu_pipelined.py: a pair is consumed
    after the next `pairs_in_flight` calls were entered, the latest moment the documented contract allows;
  * synchronous calls.
The rasterizer itself is replaced by a stub that records which buffers each eye's work touches (gsb_raster_forward_pair:
shared part on the left stream, right eye behind an event -- csrc/gsb_raster.cu)."""
import itertools

import pytest
import torch

from gs2mesh_b200 import renderer as rmod


# ------------------------------------------------------------------------------------------------ fake CUDA runtime
class Graph:
    """Happens-before graph of enqueued operations."""

    def __init__(self):
        self.deps = []      # node -> set of direct predecessors
        self.access = []    # node -> list of (resource, is_write)
        self.label = []
        self.host = None    # node every later enqueue depends on (host synchronisations)
        self.current = None  # current stream

    def node(self, label, deps=(), access=()):
        d = {x for x in deps if x is not None}
        if self.host is not None:
            d.add(self.host)
        self.deps.append(d)
        self.access.append(list(access))
        self.label.append(label)
        return len(self.deps) - 1

    def host_wait(self, node):
        """The host has waited for `node`: everything enqueued from now on comes after it."""
        if node is not None:
            self.host = self.node("host", deps=[node])

    def ancestors(self):
        anc = []
        for i, d in enumerate(self.deps):  # nodes are created in topological order
            a = set(d)
            for x in d:
                a |= anc[x]
            anc.append(a)
        return anc

    def races(self):
        anc = self.ancestors()
        touched = {}
        bad = []
        for j, acc in enumerate(self.access):
            for res, wr in acc:
                for i, wr_i in touched.get(res, []):
                    if (wr or wr_i) and i not in anc[j]:
                        bad.append((self.label[i], self.label[j], res))
                touched.setdefault(res, []).append((j, wr))
        return bad


G = None  # the graph of the running test


class FakeEvent:
    def __init__(self, enable_timing=False):
        self.node = None

    def record(self, stream=None):
        stream = stream or G.current
        self.node = stream.op("event")

    def synchronize(self):
        G.host_wait(self.node)


class FakeStream:
    _ids = itertools.count()

    def __init__(self, device=None, priority=0):
        self.id = next(FakeStream._ids)
        self.last = None
        self.pending = set()  # event nodes waited for since the last operation
        self.cuda_stream = 1000 + self.id
        self.priority = priority

    def op(self, label, access=()):
        n = G.node(f"s{self.id}:{label}", deps=[self.last, *self.pending], access=access)
        self.last = n
        self.pending = set()
        return n

    def wait_event(self, ev):
        if ev.node is not None:
            self.pending.add(ev.node)

    def record_event(self):
        ev = FakeEvent()
        ev.record(self)
        return ev

    def wait_stream(self, other):
        self.wait_event(other.record_event())

    def synchronize(self):
        G.host_wait(self.op("sync"))


class _StreamCtx:
    def __init__(self, st):
        self.st = st

    def __enter__(self):
        self.prev, G.current = G.current, self.st

    def __exit__(self, *a):
        G.current = self.prev


class FakeBuf:
    """A tensor of one buffer set; copy_ records a read of the source and a write of the destination on the current stream."""

    def __init__(self, name):
        self.name = name

    def copy_(self, src, non_blocking=False):
        G.current.op(f"copy {src.name}->{self.name}", access=[(src.name, False), (self.name, True)])
        return self

    def numpy(self):
        return None


def read_on_current(bufs, what):
    G.current.op(f"consume {what}", access=[(b.name, False) for b in bufs])


@pytest.fixture
def fake_cuda(monkeypatch):
    global G
    G = Graph()
    main = FakeStream()
    G.current = main
    monkeypatch.setattr(torch.cuda, "Stream", FakeStream)
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda device=None: G.current)
    monkeypatch.setattr(torch.cuda, "stream", _StreamCtx)
    yield main
    G = None


class _View:
    width, height = 64, 48


class _Table:
    device = "fake"


def make_renderer(n_views, in_flight, spare, copy_streams):
    r = rmod.Renderer.from_scene([dict(left={}, right={}) for _ in range(n_views)], 0.1, None)
    r.pairs_in_flight, r.spare_buffer_sets, r.copy_streams = in_flight, spare, copy_streams
    r._views = [[_View(), _View()] for _ in range(n_views)]
    r._camera_table = _Table()
    r._status = torch.zeros(n_views, 2, 4, dtype=torch.int64)
    r._shared_depth = [True] * n_views
    r._min_instances = 0
    r._ready = True
    sets = {}

    def buffers(w, h, parity=0):
        if parity not in sets:
            mk = lambda n: FakeBuf(f"set{parity}.{n}")
            sets[parity] = dict(color=[mk("color0"), mk("color1")], u8=[mk("u8_0"), mk("u8_1")], depth=mk("depth"),
                                final_T=mk("final_T"), host_u8=[mk("host_u8_0"), mk("host_u8_1")], host_depth=mk("host_depth"))
        return sets[parity]

    def enqueue_pair(camera_number, b, streams):
        # gsb_raster_forward_pair: preprocess (+ shared depth sort) on the left stream, the right eye's stream waits for an
        # event recorded behind it, then each eye bins / blends on its own stream; image_to_u8 per eye (renderer.py)
        streams[0].op("pair: shared part")
        if streams[1] is not streams[0]:
            streams[1].wait_stream(streams[0])
        streams[0].op("blend L", access=[(b["color"][0].name, True), (b["depth"].name, True), (b["final_T"].name, True)])
        streams[1].op("blend R", access=[(b["color"][1].name, True)])
        for s in range(2):
            streams[s].op(f"to_u8 {s}", access=[(b["color"][s].name, False), (b["u8"][s].name, True)])

    r._buffers = buffers
    r._enqueue_pair = enqueue_pair
    r.sets = sets
    return r


def device_consumer(out):
    """TSDF.integrate on the caller's stream: reads the left eye's depth, transmittance and uint8 frame."""
    read_on_current([out["depth"], out["final_T"], out["left_u8"]], "integrate")


def host_consumer(out):
    """bench.py's fuse(): the pinned frames are uploaded again on the caller's stream (reads of host_u8)."""
    read_on_current([out["host_left_u8"], out["host_right_u8"]], "upload frames")


CONFIGS = [(f, s, c) for f in (1, 2, 3) for s in (0, 1, 2) for c in (False, True)]


@pytest.mark.parametrize("in_flight,spare,copy_streams", CONFIGS)
def test_device_resident_loop_is_race_free(fake_cuda, in_flight, spare, copy_streams):
    r = make_renderer(4, in_flight, spare, copy_streams)
    for n in range(14):
        out = r.render_image_pair(n % 4, to_host=False)
        device_consumer(out)  # enqueued on the caller's stream, nothing waits on the host
    assert G.races() == []
    assert len(r.sets) == in_flight + 2 + spare


@pytest.mark.parametrize("order", ["lagged", "overlap", "pipelined_test"])
@pytest.mark.parametrize("in_flight,spare,copy_streams", CONFIGS)
def test_host_buffer_loops_are_race_free(fake_cuda, in_flight, spare, copy_streams, order):
    """The three consumption orders in the tree; `overlap` and `pipelined_test` touch a pair's buffers after the next
    `pairs_in_flight` calls were entered -- the last moment the contract allows."""
    r = make_renderer(4, in_flight, spare, copy_streams)
    views = [n % 4 for n in range(15)]
    lag = max(1, in_flight)
    pending = [r.render_image_pair(v, to_host=True, wait=False) for v in views[:lag]]
    fuse_next = None
    for n in range(len(views)):
        if order == "lagged":
            if fuse_next is not None:
                host_consumer(fuse_next)
            out = pending.pop(0)
            if n + lag < len(views):
                pending.append(r.render_image_pair(views[n + lag], to_host=True, wait=False))
            out["ready"].synchronize()
            device_consumer(out)  # prepare_depth on the caller's stream
            fuse_next = out
        else:
            out = pending.pop(0)
            if order == "overlap":
                out["ready"].synchronize()
                device_consumer(out)
            if n + lag < len(views):
                pending.append(r.render_image_pair(views[n + lag], to_host=True, wait=False))
            if order != "overlap":
                out["ready"].synchronize()
                device_consumer(out)
            host_consumer(out)
    if fuse_next is not None:
        host_consumer(fuse_next)
    assert G.races() == []


@pytest.mark.parametrize("in_flight,spare,copy_streams", CONFIGS)
def test_completion_handle_and_synchronous_call_cover_everything(fake_cuda, in_flight, spare, copy_streams):
    """After `ready.synchronize()` -- or after a synchronous call returns -- every operation of the pair (both blends, both
    conversions, both D2H copies) has happened for the host; after a device-path call, work enqueued on the caller's stream
    comes after both eyes."""
    r = make_renderer(4, in_flight, spare, copy_streams)
    for n, mode in enumerate(["async", "sync", "device", "async", "device", "sync", "async", "async", "sync"]):
        first = len(G.deps)
        if mode == "async":
            out = r.render_image_pair(n % 4, to_host=True, wait=False)
            mine = [i for i in range(first, len(G.deps)) if G.access[i]]
            out["ready"].synchronize()
            marker = G.node("host marker")
        elif mode == "sync":
            out = r.render_image_pair(n % 4, to_host=True)
            mine = [i for i in range(first, len(G.deps)) if G.access[i]]
            marker = G.node("host marker")
        else:
            out = r.render_image_pair(n % 4, to_host=False)
            mine = [i for i in range(first, len(G.deps)) if G.access[i]]
            marker = fake_cuda.op("caller's next kernel")
        assert len(mine) == (6 if mode != "device" else 4), (mode, [G.label[i] for i in mine])
        anc = G.ancestors()[marker]
        missing = [G.label[i] for i in mine if i not in anc]
        assert not missing, (mode, missing)
    assert G.races() == []


def test_the_detector_sees_a_missing_gate(fake_cuda, monkeypatch):
    """Self-check of the checker: with the gates of the rotation removed the same loop must be reported as racy."""
    r = make_renderer(4, 2, 0, False)
    monkeypatch.setattr(FakeStream, "wait_event", lambda self, ev: None)
    for n in range(12):
        out = r.render_image_pair(n % 4, to_host=False)
        device_consumer(out)
    assert G.races()


def test_contract_bound_is_tight(fake_cuda):
    """Consuming a pair one call LATER than the contract allows (after call n + pairs_in_flight + 1 was entered) is a race:
    the documented bound is the real one, not a conservative one."""
    in_flight = 2
    r = make_renderer(4, in_flight, 0, False)
    outs = []
    for n in range(12):
        outs.append(r.render_image_pair(n % 4, to_host=True, wait=False))
        if n >= in_flight + 1:
            late = outs[n - in_flight - 1]  # call n - 3, touched after call n (= its call + pairs_in_flight + 1) was entered
            late["ready"].synchronize()
            host_consumer(late)
    assert G.races()
