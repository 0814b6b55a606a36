import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
model_type = sys.argv[2] if len(sys.argv) > 2 else "mit_b0"
dev = torch.device("cuda", 0)
model = bench.build_model(model_type, "bf16", dev)
model.setup_runtime()
batch = bench.synth_batch(size, 2, 100, dev)
for i in range(int(sys.argv[3]) if len(sys.argv) > 3 else 1):
    model.training_step(batch, i)
    torch.cuda.synchronize()
    print("step ok", i, {k: float(v) for k, v in model._logged.items()}, flush=True)
