// launch.cuh — per-thread count of kernel launches issued by this library (reported as gpu_launches by bench.py).
#pragma once
extern thread_local unsigned g_xfg_launches;
#define XFG_LAUNCHED(k) (g_xfg_launches += (k))
