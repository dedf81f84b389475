import sys, torch
sys.path.insert(0, '/root/repo')
from jabd_b200 import anchors, batched, config, synth
VAR=(0.1,0.2)
size, B = int(sys.argv[1]), int(sys.argv[2])
widths = [int(x) for x in sys.argv[3].split(',')]
pri = anchors.Anchors(config.cfg_mnet, image_size=(size, size)).get_anchors()
ls, cs, ms = [], [], []
for i in range(B):
    gt = synth.make_gt(3, i, (size, size), count=60)
    l, c, m = synth.make_preds_clustered(3, i, pri, gt, VAR, device="cuda")
    ls.append(l.cuda()); cs.append(c.cuda()); ms.append(m.cuda())
loc, conf, landm = torch.stack(ls).contiguous(), torch.stack(cs).contiguous(), torch.stack(ms).contiguous()
ref = batched.detect(loc, conf, landm, pri, VAR, cluster=1)
torch.cuda.synchronize()
for w in widths:
    out = batched.detect(loc, conf, landm, pri, VAR, cluster=w)
    torch.cuda.synchronize()
    print("width", w, "equal", all(torch.equal(a, b) for a, b in zip(out, ref)), flush=True)
