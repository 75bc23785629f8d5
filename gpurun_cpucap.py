import time, multiprocessing as mp
def work(_):
    t=time.perf_counter(); x=0
    for i in range(3_000_000): x+=i*i
    return time.perf_counter()-t
if __name__=='__main__':
    for n in (1,8,16,32,64,128):
        t=time.perf_counter()
        with mp.Pool(n) as p: r=p.map(work, range(n))
        dt=time.perf_counter()-t
        print(n, 'procs: wall %.2fs, per-task avg %.2fs, throughput %.1f tasks/s'%(dt, sum(r)/n, n/dt))
