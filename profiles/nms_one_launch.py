import os, sys, torch
sys.path.insert(0, os.getcwd())
from jabd_b200 import _lib, _tensor, anchors, config, synth, utils_bbox
from jabd_b200._tensor import ptr
VAR=(0.1,0.2); torch.cuda.set_device(0); dev=torch.device("cuda",0); L=_lib.lib()
size,B=640,32
pri=anchors.Anchors(config.cfg_mnet,image_size=(size,size)).get_anchors(); P=pri.shape[0]
boxes,scores=[],[]
for i in range(B):
    gt=synth.make_gt(3,i,(size,size),count=60)
    l,c,m=synth.make_preds_clustered(3,i,pri,gt,VAR,device="cuda")
    boxes.append(utils_bbox.decode(l.cuda(),pri,VAR)); scores.append(c.cuda()[:,1].contiguous())
bx,sc=torch.stack(boxes).contiguous(),torch.stack(scores).contiguous()
keep=torch.empty((B,750),dtype=torch.int32,device=dev); cnt=torch.empty((B,),dtype=torch.int32,device=dev)
ws=_tensor.workspace(L.jabd_nms_workspace_bytes(B,P,750),dev)
width=int(sys.argv[1]) if len(sys.argv)>1 else 0
def run(): _lib.call("jabd_nms",ptr(bx),P*4,4,ptr(sc),P,1,B,P,0.02,2,5000,0.4,(width<<12),750,ptr(keep),ptr(cnt),ptr(ws),ws.numel(),_tensor.stream_of(dev))
for _ in range(3): run()
torch.cuda.synchronize()
torch.cuda.profiler.start(); run(); torch.cuda.synchronize(); torch.cuda.profiler.stop()
print("done", cnt.float().mean().item())
