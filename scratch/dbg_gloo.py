import torch.multiprocessing as mp, sys
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import test_dist_gloo as t
if __name__=="__main__":
    ctx=mp.get_context("spawn"); q=ctx.Queue(); port=t._free_port()
    ps=[ctx.Process(target=t._worker,args=(r,2,port,q)) for r in range(2)]
    [p.start() for p in ps]
    [p.join(60) for p in ps]
    while not q.empty(): print(q.get())
