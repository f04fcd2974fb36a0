import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import runpy
for flag in (False, True):
    torch.backends.cudnn.benchmark = flag
    print("cudnn.benchmark =", flag, flush=True)
    sys.argv = ["bench_decoder.py", "8"]
    runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), "bench_decoder.py"), run_name="__main__")
