"""Print the error margins of the default (tcgen05) decoder path against the
golden vectors of the unmodified reference (GPU only; diagnostics)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import pyroved_b200 as pv  # noqa: E402
from golden_util import Golden  # noqa: E402

for generic in ("1", "0"):
    os.environ["PVB_FORCE_GENERIC"] = generic
    for name in ["ivae_1d_t", "ivae_28_rt", "ivae_28_rt_beta3"]:
        g = Golden(name)
        m = pv.models.iVAE(seed=1, device="cuda:0", **g.kwargs)
        m.load_state_dict(g.group("w0"))
        tr = pv.trainers.SVItrainer(m, device="cuda:0")
        x, _ = g.args()
        kw = {k: float(v) for k, v in g.kw().items()}
        loss = tr.svi.loss_and_grads(x.cuda(), _eps=g.eps().cuda(), **kw)
        prog = next(iter(tr.svi.programs.values()))
        loc_err = (prog.loc.cpu().reshape(-1) - g.t("loc").reshape(-1)).abs().max().item()
        worst = ("", 0.0)
        for k, p in m.named_parameters():
            ref = g.group("grad")[k].cuda()
            e = (p.grad - ref).abs().max().item() / (ref.abs().max().item() + 1e-6)
            if e > worst[1]:
                worst = (k, e)
        print("{:18s} path={:8s} loss rel {:.2e}  loc max|err| {:.2e}  worst grad rel {:.2e} ({})".format(
            name, "generic" if generic == "1" else "tcgen05", abs(loss - g.loss) / abs(g.loss),
            loc_err, worst[1], worst[0]))
