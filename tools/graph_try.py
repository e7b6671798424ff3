import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch, dxrvoxelizer_b200 as d
m = d.load_obj(d.asset_path("dragon.obj"))
s = torch.cuda.Stream()
vox = d.Voxelizer(0); vox.set_stream(s.cuda_stream)
vb = torch.from_numpy(m.vertex_bytes).cuda(); ib = torch.from_numpy(m.indices.view(np.int32)).cuda()
def step():
    vox.build_bvh_device(vb.data_ptr(), m.num_vertices, m.stride, ib.data_ptr(), m.indices.size)
    vox.voxelize(1024, d.MODE_PARITY)
for _ in range(3): step()
torch.cuda.synchronize()
def timeit(fn, iters=50):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(iters): fn()
    e1.record(s); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/iters*1e3
print("stream launches: %.1f us/step" % timeit(step))
g = torch.cuda.CUDAGraph()
with torch.cuda.stream(s):
    g.capture_begin()
    step()
    g.capture_end()
torch.cuda.synchronize()
with torch.cuda.stream(s):
    print("graph replay:    %.1f us/step" % timeit(g.replay))
ref = vox.count_inside()
print("inside", ref)
