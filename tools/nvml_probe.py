import time, threading, torch, pynvml
pynvml.nvmlInit(); h=pynvml.nvmlDeviceGetHandleByIndex(0)
x=torch.randn(4096,4096,device='cuda'); 
def work(n):
    ts=[]
    for _ in range(n):
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200): y=x*1.0001
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return ts
work(3)
print('baseline steps ms', [round(t,2) for t in work(10)])
for name,fn in [('clock',lambda: pynvml.nvmlDeviceGetClockInfo(h,pynvml.NVML_CLOCK_SM)),
                ('reasons',lambda: pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)),
                ('power',lambda: pynvml.nvmlDeviceGetPowerUsage(h))]:
    lat=[]
    stop=False
    def poll():
        while not stop:
            t=time.perf_counter(); fn(); lat.append((time.perf_counter()-t)*1e3); time.sleep(0.05)
    th=threading.Thread(target=poll); th.start()
    ts=work(10); stop=True; th.join()
    print(name,'call ms',[round(l,2) for l in lat[:12]],'steps ms',[round(t,2) for t in ts])
