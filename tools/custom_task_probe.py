"""Probe of bench.py's custom-task leg alone (recorded step / two-launch / generic), K steps each."""
import json
import sys

import torch as th

sys.path.insert(0, ".")
import bench  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 400
dev = th.device("cuda", 0)
th.cuda.set_device(dev)
stream = th.cuda.current_stream(dev)
out = bench.custom_task_leg(bench.AGENTS, dev, K, 30, stream, th.cuda.synchronize)
print(json.dumps(out, indent=1))
