"""Where the batched segment classifier (WindowedSqueezeNet, library kernels) spends its time.

    python profiles/classifier_timing.py [benchmark]
"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from swiftwatcher_b200 import segment_classification as sc

torch.backends.cudnn.benchmark = "benchmark" in sys.argv
torch.manual_seed(0)
model = sc.setup_model(2, torch.device("cuda:0"))
clf = sc.SegmentClassifier(model.state_dict(), device="cuda:0", channels_last="nchw" not in sys.argv)
crops = torch.randint(0, 256, (8192, 24, 24, 3), dtype=torch.uint8, device="cuda:0")
for _ in range(3):
    clf.scores(crops)
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(5):
    clf.scores(crops)
torch.cuda.synchronize()
dt = (time.perf_counter() - t) / 5
print("channels_last=%s cudnn.benchmark=%s tf32=%s: %.2f ms per 8192 crops = %.0f crops/s" % (clf.channels_last, torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, dt * 1e3, 8192 / dt))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    clf.scores(crops)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))
