import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
import xfg_stark_b200 as xs, orc
n_log2 = int(sys.argv[1]); o = tuple(int(v) for v in sys.argv[2].split(","))
s = orc.synthetic_inputs(0)
air = xs.pack_inputs(s["burn"], s["mint"], s["tx_prefix_hash"], s["recipient"], s["secret"], s["network_id"], s["target_chain_id"], s["version"])
trace = xs.build_trace(air, n_log2)
with xs.Context(device=0, max_n_log2=min(24, n_log2 + 2), num_slots=1) as ctx:
    d = torch.from_numpy(np.ascontiguousarray(trace).view(np.int64)).cuda()
    for _ in range(2):
        p, t = ctx.prove_device(d.data_ptr(), n_log2, air, xs.ProofOptions(*o), want_times=True)
    print(len(p), t["device_ms"], t["kernel_launches"])
