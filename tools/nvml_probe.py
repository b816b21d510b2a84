import time, pynvml, torch, threading
pynvml.nvmlInit(); dev=pynvml.nvmlDeviceGetHandleByIndex(0)
x=torch.zeros(1<<20,device='cuda')
def work(n):
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(n):
        x.add_(1.0)
    torch.cuda.synchronize(); return (time.perf_counter()-t)/n*1e6
print("launch us/iter baseline", work(20000))
calls={"clock":lambda: pynvml.nvmlDeviceGetClockInfo(dev,pynvml.NVML_CLOCK_SM),
       "reasons":lambda: pynvml.nvmlDeviceGetCurrentClocksEventReasons(dev),
       "power":lambda: pynvml.nvmlDeviceGetPowerUsage(dev),
       "maxclock":lambda: pynvml.nvmlDeviceGetMaxClockInfo(dev,pynvml.NVML_CLOCK_SM)}
for name,fn in calls.items():
    t=time.perf_counter(); 
    for _ in range(20): fn()
    print(name, "call ms", (time.perf_counter()-t)/20*1e3)
for name,fn in calls.items():
    stop=threading.Event()
    def loop():
        while not stop.is_set():
            fn(); stop.wait(0.05)
    th=threading.Thread(target=loop); th.start()
    print("launch us/iter with", name, "every 50ms:", work(20000))
    stop.set(); th.join()
