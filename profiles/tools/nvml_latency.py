import os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import pynvml as n
import circuitsimulator_b200 as bg, bench
cfg, Gd, Hd, samples, k, desc = bench.load_config("hidden_shift_n40_t40_k9_L8192")
t = cfg["t"]; L = bench.fixed_L(k, t)
G = bg.Projector.make(*Gd); H = bg.Projector.make(*Hd)
ctx = bg.Backend(0); ctx.set_decomposition(t, 0, L); ctx.sampled_prepare2(G, H, samples, 1, 1, 2)
stop = False; steps = [0]; gaps = []
def load():
    last = time.perf_counter()
    ctx.sampled_run()
    while not stop:
        ctx.sampled_run(); ctx.sampled_finish2(1.0); steps[0] += 1
        now = time.perf_counter(); gaps.append(now - last); last = now
    ctx.sampled_finish2(1.0)
th = threading.Thread(target=load); th.start()
time.sleep(0.5)
n.nvmlInit(); h = n.nvmlDeviceGetHandleByIndex(0)
calls = {"clock": lambda: n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM), "maxclock": lambda: n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM),
         "power": lambda: n.nvmlDeviceGetPowerUsage(h), "reasons": lambda: n.nvmlDeviceGetCurrentClocksEventReasons(h)}
for name, f in calls.items():
    ts = []
    for _ in range(30):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0); time.sleep(0.01)
    ts.sort(); print(name, "median %.3f ms max %.3f ms" % (ts[15] * 1e3, ts[-1] * 1e3), flush=True)
g0 = len(gaps); time.sleep(0.5); base = sorted(gaps[g0:]); 
print("step gap median %.3f ms p99 %.3f max %.3f (no nvml)" % (base[len(base)//2]*1e3, base[int(len(base)*0.99)]*1e3, base[-1]*1e3))
g0 = len(gaps)
for _ in range(10):
    for f in calls.values(): f()
    time.sleep(0.05)
w = sorted(gaps[g0:]); print("step gap median %.3f ms p99 %.3f max %.3f (with nvml polls)" % (w[len(w)//2]*1e3, w[int(len(w)*0.99)]*1e3, w[-1]*1e3))
stop = True; th.join(); ctx.close()
