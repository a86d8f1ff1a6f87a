import os, sys, time, subprocess, threading
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from comet_b200 import capi
import pynvml
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
L = capi.lib()
dev = torch.device("cuda", 0)
n, d, K, nq = 1_000_000, 768, 100, 512
g = torch.Generator(device=dev); g.manual_seed(1)
x = torch.randn((n, d), generator=g, device=dev)
ix = capi.FlatIndex(d, capi.COSINE)
ix.add_device(np.arange(1, n + 1, dtype=np.uint32), x.data_ptr(), n)
del x
q = torch.randn((nq, d), generator=g, device=dev)
oi = torch.zeros((nq, K), dtype=torch.int32, device=dev); osc = torch.zeros((nq, K), device=dev); oc = torch.zeros(nq, dtype=torch.int64, device=dev)
st = torch.cuda.current_stream().cuda_stream
def run(tag, secs=2.0):
    samples = []; pw = []
    stop = threading.Event()
    def samp():
        while not stop.is_set():
            samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)); pw.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
            time.sleep(0.02)
    for _ in range(3):
        ix.search_device(q.data_ptr(), nq, K, oi.data_ptr(), osc.data_ptr(), oc.data_ptr(), K, stream=st, path=capi.PATH_TENSOR)
    torch.cuda.synchronize()
    t = threading.Thread(target=samp); t.start()
    t0 = time.time(); reps = 0
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < secs:
        for _ in range(50):
            ix.search_device(q.data_ptr(), nq, K, oi.data_ptr(), osc.data_ptr(), oc.data_ptr(), K, stream=st, path=capi.PATH_TENSOR)
        reps += 50
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    stop.set(); t.join()
    ms = e0.elapsed_time(e1) / reps
    half = samples[len(samples)//2:]; ph = pw[len(pw)//2:]
    print(f"{tag}: {ms:.3f} ms/step  sm clock median {np.median(half):.0f} MHz (min {min(half)}), power {np.median(ph):.0f} W", flush=True)
for kv in sys.argv[1:]:
    k, v = kv.split("=")
    os.environ[k] = v
    run(kv)
