import importlib, sys
import numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle'); sys.path.insert(0,'/root/repo/tests')
rtb = importlib.import_module("ray-tracing-v06_b200"); import pyoracle as orc
from helpers import *
r=rtb.Renderer(0)
for name,lo,hi in (("book2_cornell",0,555),("book2_final",-200,600)):
    scene=rtb.Scene.named(name); r.set_scene(scene)
    rays=random_rays(rtb,100000,lo,hi,seed=11)
    g=r.trace_rays(rays); o=orc.OracleScene(scene.serialize()).trace_rays(rays,rtb.HIT_DTYPE)
    hit=(g["object"]>=0)&(g["object"]==o["object"])
    bad=hit&(g["n"].view(np.uint32)!=o["n"].view(np.uint32)).any(axis=1)
    print(name,"bad",bad.sum(),"of",hit.sum())
    h,mats,objs,ch=parse_blob(scene.serialize())
    idx=np.nonzero(bad)[0][:6]
    for i in idx:
        print(" obj",g["object"][i],"kind",objs[g["object"][i]]["kind"],"gn",g["n"][i],"on",o["n"][i],"ff",g["front_face"][i],o["front_face"][i], "t",g["t"][i],o["t"][i])
    kinds=set(int(objs[k]["kind"]) for k in g["object"][bad]); print(" kinds with bad n:",kinds)
