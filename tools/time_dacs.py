"""How long does the eager DACS mix between the two CUDA graphs of the step take (host + device)?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
model = bench.build_model("mit_b5", "bf16", dev)
model.setup_runtime()
batch = bench.synth_batch(1024, 2, 100, dev)
model.enable_cuda_graphs(warmup=1)
for i in range(5):
    model.training_step(batch, i)
torch.cuda.synchronize()
G = model._graphs
out, sb = G['out_a'], G['batch']
ts = []
for i in range(20):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    m = model.get_dacs_mix(out['images_trg'], out['probs'], sb['image_src'], sb['semantic_src'], fused=out['fused'])
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    ts.append(((t1 - t0) * 1e3, (t2 - t0) * 1e3))
ts.sort(key=lambda t: t[1])
print("dacs host-issue ms / total ms (median):", ts[len(ts) // 2])
for name, g in (("A", G['a']), ("B", G['b'])):
    torch.cuda.synchronize(); t0 = time.perf_counter(); g.replay(); torch.cuda.synchronize()
    print("graph", name, "ms:", (time.perf_counter() - t0) * 1e3)
