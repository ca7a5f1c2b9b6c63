import sys, os, tempfile, torch
ROOT="/root/repo"; PKG=os.path.join(ROOT,"mit-driverless-cv-traininginfra_b200")
for p in (ROOT, os.path.join(ROOT,"tests"), PKG, os.path.join(PKG,"CVC-YOLOv3"), os.path.join(PKG,"RektNet")): sys.path.insert(0,p)
import helpers
from oracle import yolo_oracle as YO
gold = torch.load(os.path.join(ROOT,"tests/golden/yolo_golden.pt"), weights_only=False)
d = tempfile.mkdtemp()
for name in ["tiny_128","tiny_416","full_128","tiny_128_c80"]:
    g = gold["darknet"][name]
    model,_ = helpers.make_darknet(d, g["cfg"], g["S"], g["C"])
    model = model.cuda().train()
    x = YO.synth_images(g["B"], g["S"], g["S"], seed=0).cuda(); tg = YO.synth_targets(g["B"],16,seed=1).cuda()
    with torch.no_grad(): out = model(x,tg)
    got = torch.stack([o.detach() for o in out]).cpu()
    rel = (got-g["losses"]).abs()/g["losses"].abs().clamp_min(1e-3)
    print(name, ["%.1e"%v for v in rel.tolist()])
