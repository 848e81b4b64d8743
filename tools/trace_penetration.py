"""Per-warp timeline of the two penetration-depth stages of one 1 M-configuration pass (scenario 011), from the -DJG_TRACE
build of the library (build it here first: python -c "from jogramop_framework_b200 import build; build.build_trace()").
Never a timed or shipped path.  Output of the round-2 run: profiles/r02_penetration_timeline.txt."""
import sys, os, ctypes as C, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.chdir(ROOT)
os.environ.setdefault('JG_B200_LIB_PATH', os.path.join(ROOT, 'jogramop_framework_b200', 'libjogramop_b200_trace.so'))
from jogramop_framework_b200 import _lib
from jogramop_framework_b200.engine import CollisionEngine
from jogramop_framework_b200.ingest import RobotModel, SceneModel
D='jogramop_framework_b200/data'
robot=RobotModel.load(f'{D}/robot_mobile_panda.npz'); scene=SceneModel.load(f'{D}/scene_011.npz')
lim=robot.joint_limits[robot.arm_joint_ids]
q=torch.from_numpy((lim[:,0]+(lim[:,1]-lim[:,0])*np.random.default_rng(11).random((1000000,9))).astype(np.float32)).cuda()
e=CollisionEngine(robot,scene)
for _ in range(4): e.check(q,flags=3,packed=True)
torch.cuda.synchronize()
L=_lib.load(); L.jg_trace_read.argtypes=[C.c_void_p]
buf=np.zeros((2,4096,8),np.uint64)
assert L.jg_trace_read(buf.ctypes.data_as(C.c_void_p))==0
for s in range(2):
    r=buf[s]; r=r[r[:,1]>0]
    t0=r[:,0].min(); dur=(r[:,1].astype(np.int64)-int(t0))/1e3; start=(r[:,0].astype(np.int64)-int(t0))/1e3
    print('launch slot',s,'warps',len(r),'kernel span us %.1f'%dur.max(),'warp end us: p10 %.1f p50 %.1f p90 %.1f p99 %.1f'%tuple(np.percentile(dur,[10,50,90,99])),'start max %.1f'%start.max())
    print('   keys visited: mean %.1f max %d; polytopes started per warp: mean %.1f min %d max %d'%(r[:,2].mean(),r[:,2].max(),r[:,3].mean(),r[:,3].min(),r[:,3].max()))
    # per block (13 warps): block end = max
    nb=len(r)//13
    be=dur[:nb*13].reshape(nb,13).max(1)
    print('   block end us: min %.1f p10 %.1f p50 %.1f p90 %.1f max %.1f'%(be.min(),*np.percentile(be,[10,50,90]),be.max()))
    last=(r[:,4].astype(np.int64)-int(t0))/1e3
    print('   last polytope start us: p10 %.1f p50 %.1f p90 %.1f max %.1f ; drain (end-last start) us: p50 %.1f p90 %.1f max %.1f'%(*np.percentile(last,[10,50,90]),last.max(),*np.percentile(dur-last,[50,90]),(dur-last).max()))
    print('   steps per warp: mean %.1f min %d max %d ; steps after last start: p50 %d p90 %d max %d ; us/step mean %.2f'%(r[:,6].mean(),r[:,6].min(),r[:,6].max(),*np.percentile(r[:,7],[50,90]),r[:,7].max(),(dur/np.maximum(r[:,6],1)).mean()))
    late=np.argsort(dur)[-8:]
    print('   latest: (warp, end, lastStart, key0, started, steps, stepsAfterLast)',[(int(i),round(float(dur[i]),1),round(float(last[i]),1),int(r[i,5]),int(r[i,3]),int(r[i,6]),int(r[i,7])) for i in late])
    early=np.argsort(dur)[:5]
    print('   earliest:',[(int(i),round(float(dur[i]),1),round(float(last[i]),1),int(r[i,5]),int(r[i,3]),int(r[i,6]),int(r[i,7])) for i in early])
    late=np.argsort(dur)[-5:]
    print('   latest warps:',[(int(i),round(float(dur[i]),1),int(r[i,2]),int(r[i,3])) for i in late])
