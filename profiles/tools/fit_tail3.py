"""Tail slowdown of Model.fit: (1) as is, (2) final wait by polling, (3) producer thread lingers before exiting, (4) both."""
import os, sys, time, queue as _queue, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
from handyrec_b200 import lowering
vocabs = bench.CRITEO_VOCABS
model, KL = bench.build_deepfm_model(vocabs)
model.compile(optimizer=KL.Adam(learning_rate=1e-3), loss=KL.binary_crossentropy)
B, steps = bench.BATCH, 20
g = np.random.RandomState(0)
x = {f"C{i + 1}": g.randint(0, v, (B * steps, 1)).astype(np.int32) for i, v in enumerate(vocabs)}
x.update({f"I{j + 1}": g.rand(B * steps, 1).astype(np.float32) for j in range(bench.N_DENSE)})
y = (g.rand(B * steps) < 0.25).astype(np.float32)
model.fit({k: v[: 10 * B] for k, v in x.items()}, y[: 10 * B], batch_size=B, epochs=1)
torch.cuda.synchronize()
eng = model._fused.engine
orig = eng.train_step_on_device
ev = []


def wrapped(*a):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    orig(*a)
    e1.record()
    ev.append((e0, e1))


eng.train_step_on_device = wrapped
real_sync = torch.cuda.Stream.synchronize


def poll_sync(self):
    while not self.query():
        pass


class LingerQueue(_queue.Queue):
    def put(self, item, *a, **k):
        super().put(item, *a, **k)
        if item is None:
            time.sleep(0.03)


shim = types.SimpleNamespace(Queue=LingerQueue)


def run(tag, poll, linger):
    torch.cuda.Stream.synchronize = poll_sync if poll else real_sync
    lowering.queue = shim if linger else _queue
    for _ in range(3):
        ev.clear()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        model.fit(x, y, batch_size=B, epochs=1)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        d = [a.elapsed_time(b) for a, b in ev]
        print(f"{tag}: wall {1e3 * (t1 - t0):.2f} ms ({1e3 * (t1 - t0) / len(d):.3f}/step); tail " + " ".join(f"{v:.2f}" for v in d[-6:]), flush=True)
    torch.cuda.Stream.synchronize = real_sync
    lowering.queue = _queue


from handyrec_b200 import _lib
for on in (1, 0, 1, 0):
    kind = _lib.lib().hrb_host_pack_set_writeback(on)
    run(f"cache write-back of packed blocks {'on' if on else 'off'} (cpu offers {kind})", False, False)
