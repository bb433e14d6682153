import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cts_b200 as cts
import bench
ctx = cts.Context.default()
stream = torch.cuda.current_stream()
for dt in (np.float64, np.float32):
    r = bench.run_c4(cts, ctx, stream, dt)
    print(json.dumps(r))
