"""Host-only statistics of the tile layout on the Kuhn grid (no GPU): tets / vertices per tile, phase C rows per warp,
row fill, part C bytes per tet, staging-slot usage.  The numbers quoted in DESIGN.md section 7 ("What actually bounds
k_local") come from `python scripts/layout_stats.py 55`."""
import importlib, sys, numpy as np, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
pd = importlib.import_module("soft-body-simulation-cuda_b200")
n=int(sys.argv[1]) if len(sys.argv)>1 else 55
t=time.time()
sc = pd.Scene.kuhn_grid(n,n,n,1.0,0.05,12345,(0,10,0),1.0,2e5)
L = sc.layout()
print("layout", time.time()-t, "tiles", L.num_tiles, "slots", L.num_slots)
T = L.tile_table
ntet = T[:,2] & 0xffff; nloc = T[:,2] >> 16
g = T[:,4:12]
nR = (g>>6)&63; nVg=(g>>12)&63
print("tets/tile mean", ntet.mean(), "min", ntet.min(), "verts/tile mean", nloc.mean(), "max", nloc.max())
print("rows per group idx mean:", nR.mean(0).round(2))
print("verts per group idx mean:", nVg.mean(0).round(1))
print("rows total per tile mean", nR.sum(1).mean(), "max-group rows mean", nR.max(1).mean(), " sum/8:", nR.sum(1).mean()/8)
cb = T[:,1]>>16
print("part C bytes/tile mean", cb.mean(), "=> per tet", cb.sum()/ntet.sum())
# vstage padding
vs = L.vstage.reshape(-1,256)
used = (vs!=0xffffffff)
last = np.where(used.any(1), 255-np.argmax(used[:,::-1],axis=1), 0)
print("vstage: used/tile", used.sum(1).mean(), "highest used slot mean", last.mean(), "max", last.max())
# list lengths
vl = L.vlist.reshape(-1,256)
# incidence count per local vertex = from rows? approximate by total entries 4*ntet
print("entries/tile", (4*ntet).mean(), "rows*64 capacity", (nR.sum(1)*64).mean(), "fill", (4*ntet).sum()/(nR.sum(1)*64).sum())
# per-group fill: count real entries per group from part C rows
rec = L.records
tot_real = np.zeros(8); tot_cap = np.zeros(8)
ZERO = 4*256*16
for ti in range(0, L.num_tiles, 7):
    off = int(T[ti,0])*16; ab = int(T[ti,1]&0xffff); cb_ = int(T[ti,1]>>16)
    C = np.frombuffer(rec[off+ab:off+ab+cb_].tobytes(), np.uint16).reshape(-1,64)   # rows x (32 lanes x 2 entries)
    for gi in range(8):
        w = int(T[ti,4+gi]); rb = w&63; nr=(w>>6)&63
        if nr==0: continue
        rows = C[rb:rb+nr]
        tot_real[gi] += (rows < ZERO).sum(); tot_cap[gi] += rows.size
print("fill per group:", (tot_real/np.maximum(tot_cap,1)).round(3), "overall", tot_real.sum()/tot_cap.sum())
