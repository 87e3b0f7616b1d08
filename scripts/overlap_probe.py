import sys, time, torch
sys.path.insert(0, '/root/repo')
import bench
from toothgroupnetwork_b200 import pointnet2_utils as pn2
dev = torch.device('cuda', 0)
sa = bench.build_module(dev)
B = 592
feats = bench.make_clouds(0, B).to(dev)
xyz = feats[:, :3].contiguous()
def run(chunks, streams):
    ss = [torch.cuda.Stream() for _ in range(streams)]
    cs = B // chunks
    def step():
        main = torch.cuda.current_stream()
        for s in ss: s.wait_stream(main)
        for k in range(chunks):
            with torch.cuda.stream(ss[k % streams]):
                sa(xyz[k*cs:(k+1)*cs], feats[k*cs:(k+1)*cs])
        for s in ss: main.wait_stream(s)
    with torch.no_grad():
        for _ in range(3): step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): step()
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"chunks={chunks} streams={streams} ms/step={ms:.3f} value={B*1024/ms*1e3:.3e}")
for c, s in [(1,1),(2,2),(4,2),(4,4),(8,4),(8,8),(16,8)]:
    run(c, s)
