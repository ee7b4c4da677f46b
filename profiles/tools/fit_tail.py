"""Why are the last 3-4 steps of a fit() slower on the GPU?  Variants of the host loop, GPU step durations of the tail."""
import os, sys, time, threading, queue
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
from handyrec_b200 import lowering
vocabs = bench.CRITEO_VOCABS
model, KL = bench.build_deepfm_model(vocabs)
model.compile(optimizer=KL.Adam(learning_rate=1e-3), loss=KL.binary_crossentropy)
B, steps = bench.BATCH, 24
g = np.random.RandomState(0)
x = {f"C{i + 1}": g.randint(0, v, (B * steps, 1)).astype(np.int32) for i, v in enumerate(vocabs)}
x.update({f"I{j + 1}": g.rand(B * steps, 1).astype(np.float32) for j in range(bench.N_DENSE)})
y = (g.rand(B * steps) < 0.25).astype(np.float32)
model.fit({k: v[: 10 * B] for k, v in x.items()}, y[: 10 * B], batch_size=B, epochs=1)
torch.cuda.synchronize()
fused = model._fused
eng = fused.engine
orig = eng.train_step_on_device
ev = []


def wrapped(*a):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    orig(*a)
    e1.record()
    ev.append((e0, e1))


eng.train_step_on_device = wrapped


def report(tag, t0, t1):
    d = [a.elapsed_time(b) for a, b in ev]
    print(f"{tag}: wall {1e3 * (t1 - t0):.2f} ms ({1e3 * (t1 - t0) / len(d):.3f}/step); gpu step ms: median {sorted(d)[len(d) // 2]:.3f}; tail " + " ".join(f"{v:.2f}" for v in d[-6:]), flush=True)


def run(tag, fn):
    for _ in range(2):
        ev.clear()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        report(tag, t0, time.perf_counter())


run("Model.fit", lambda: model.fit(x, y, batch_size=B, epochs=1))

# prepacked pinned batches, no producer thread
slots = [lowering.HostBatch(eng.B, fused.ids_cols, fused.n_dense) for _ in range(steps)]
idc, dc, n = fused._columns(x)
lc = lowering._ColumnSet([y.reshape(-1, 1)])
for k, s in enumerate(slots):
    fused._pack(s, idc, dc, lc, k * B, B)
run("fit_batches(prepacked, no thread)", lambda: eng.fit_batches(iter(slots)))

# producer thread that only hands over prepacked slots
def threaded(linger):
    def gen():
        q = queue.Queue(maxsize=2)
        def prod():
            for s in slots:
                q.put(s)
            q.put(None)
            time.sleep(linger)
        th = threading.Thread(target=prod, daemon=True)
        th.start()
        while True:
            it = q.get()
            if it is None:
                break
            yield it
    return gen()
run("fit_batches(prepacked, thread)", lambda: eng.fit_batches(threaded(0.0)))
run("fit_batches(prepacked, thread lingers 30 ms)", lambda: eng.fit_batches(threaded(0.03)))

# device-resident loop for reference, same wrapper
ids_d, dense_d, label_d = eng._stage[0]
def dev_loop():
    for _ in range(steps):
        eng.train_step_on_device(ids_d[:B], dense_d[:B], label_d[:B])
run("device-resident loop", dev_loop)
def dev_loop_loss():
    for _ in range(steps):
        eng.train_step_on_device(ids_d[:B], dense_d[:B], label_d[:B])
        eng.loss_host.copy_(eng.loss_sum, non_blocking=True)
run("device-resident loop + loss D2H", dev_loop_loss)
