import sys, time, numpy as np
sys.path.insert(0,'/root/repo')
import astar_planners_b200 as b200
from astar_planners_b200 import synth
c=synth.CONFIGS['C2']
gm=b200.GridMap(c['size'],c['resolution'],origin=c['origin'],local_update_range=(60,60,20),ground_height=c['origin'][2])
gm.cloudCallback(synth.pack_cloud(synth.config_cloud('C2',n_points=1_000_000)),(0,0,1))
s,g=synth.random_queries(6,gm.getInflateOccupancy,c['origin'],c['size'],2048,min_dist=5.0,z_range=(0.3,3.0))
for algo,sem in [('astar','faithful'),('thetastar','faithful'),('astar','canonical')]:
    pl=b200.Planner(algo,resolution=0.1,lambda_heuristic=5.0,allocate_num=1000000,semantics=sem); pl.setGridMap(gm)
    pl.findPaths(s[:64],g[:64])
    t=time.time(); r,_=pl.findPaths(s,g); dt=time.time()-t
    tu=r['time_us']; it=r['iter_num']
    per=tu/np.maximum(it,1)
    print(algo,sem,'wall',round(dt,3),'sum time_us',round(tu.sum()/1e6,3),'max',round(tu.max()/1e6,3),'median us/iter',round(np.median(per),1),'p99 us/iter',round(np.percentile(per,99),1),'iter of slowest',it[tu.argmax()],'peak',r['open_peak'][tu.argmax()],'launches',pl.last_launches(), 'status', np.bincount(r['status']))
