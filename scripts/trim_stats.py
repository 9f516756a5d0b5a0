"""Host-only: tets and tiles every rank evaluates with and without PD_DIST_TRIM (csrc/layout.hpp) on the n^3-cell Kuhn grid.
    python scripts/trim_stats.py 139 8      # the numbers quoted in DESIGN.md section 9 and profiles/r1_host_checks_last_session.txt"""
import importlib, sys, os, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pd = importlib.import_module("soft-body-simulation-cuda_b200")
n=int(sys.argv[1]); worlds=[int(x) for x in sys.argv[2].split(',')]
t=time.time()
sc = pd.Scene.kuhn_grid(n,n,n,1.0,0.05,12345,(0,10,0),1.0,2e5)
G = sc.layout(); nT=sc.counts()[1]
print("global layout %.0fs tiles %d"%(time.time()-t, G.num_tiles))
for w in worlds:
    for trim in ("0","1"):
        os.environ["PD_DIST_TRIM"]=trim
        tets=[]; tiles=[]
        for r in range(w):
            P=pd.RankPlan(G,w,r); L=P.local_layout(G)
            tets.append(L.num_tets); tiles.append(L.num_tiles)
        tets=np.array(tets); tiles=np.array(tiles)
        print("world",w,"trim",trim,"tets/rank rel",(tets/(nT/w)).round(3),"max %.3f mean %.3f"%(tets.max()/(nT/w), tets.mean()/(nT/w)),"tiles/rank max",tiles.max(),"ideal",G.num_tiles//w, "%.0fs"%(time.time()-t))
