"""Which tensor methods do the UNCHANGED reference scripts call on tensors that live on the GPU when gpu > 0?

Scans the reference's Lua sources (in this container: /root/reference) and writes tests/golden/lua_tensor_methods.json:
every `receiver:method(` / `receiver[{...}]:method(` / `receiver[{...}] = ` whose receiver is a GPU tensor of that script.
GPU tensors = variables the script moves with `x = x:cuda()` plus the values derived from them or from GPU modules
(EXTRA below, each with the line that creates it).  tests/test_abi.py checks that lua/cenn.lua implements every method
listed here, plus the ones optim.adam and nn's containers call.

    python tools/scan_lua_methods.py [/root/reference]
"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = ["train.lua", "train_vid_weighted.lua", "train_deepernet.lua", "train_mycrit.lua", "test_vid_wholeim.lua", "test_vid.lua", "test.lua", "demo.lua", "inpaint_utils.lua"]
# derived GPU values: name -> where it comes from (train.lua line numbers; the video scripts repeat them)
EXTRA = {
    "fake": "netG:forward (train.lua:328)", "output": "netD:forward (train.lua:303)", "df_do": "criterion:backward (train.lua:306)",
    "df_dg": "netD:updateGradInput (train.lua:373)", "df_dg_l2": "criterionMSE:backward (train.lua:379)", "df_dg_gdl": "criterionMSE:backward (train_vid_weighted.lua:525)",
    "wtl2Matrix": "df_dg_l2:clone() (train.lua:389)", "parametersD": "netD:getParameters after :cuda() (train.lua:262)", "parametersG": "train.lua:263",
    "gradParametersD": "train.lua:262", "gradParametersG": "train.lua:263", "pred_center": "net:forward (test.lua:92)", "preds": "net:forward",
    "weights": "input_mask (train_vid_weighted.lua:494)", "mid_output": "netI:forward (train_vid_weighted.lua:402)",
    # inpaint_utils.fillIn is called with GPU tensors at train_vid_weighted.lua:404 / test_vid_wholeim.lua:188
    "dst": "inpaint_utils.fillIn argument", "src": "inpaint_utils.fillIn argument", "mask": "inpaint_utils.fillIn argument", "selection": "src:maskedSelect (inpaint_utils.lua:84)",
}
SKIP_METHODS = {"cuda"}      # the move itself


def scan(ref):
    out = []
    for fn in FILES:
        path = os.path.join(ref, fn)
        if not os.path.exists(path):
            continue
        lines = open(path, encoding="utf-8", errors="replace").read().splitlines()
        gpu = set(EXTRA)
        for ln in lines:
            code = ln.split("--")[0]
            for m in re.finditer(r"\b(\w+)\s*=\s*(\w+):cuda\(\)", code):
                if m.group(1) == m.group(2):
                    gpu.add(m.group(1))
        pat = re.compile(r"\b(\w+)(\[\{.*?\}\])?:(\w+)\(")
        for i, ln in enumerate(lines, 1):
            code = ln.split("--")[0]
            for m in pat.finditer(code):
                recv, idx, meth = m.group(1), m.group(2), m.group(3)
                if recv in gpu and meth not in SKIP_METHODS and not (fn != "inpaint_utils.lua" and recv in ("dst", "src", "mask", "selection")):
                    out.append({"file": fn, "line": i, "receiver": recv, "indexed": bool(idx), "method": meth})
            for m in re.finditer(r"\b(\w+)\[\{.*?\}\]\s*=[^=]", code):
                if m.group(1) in gpu:
                    out.append({"file": fn, "line": i, "receiver": m.group(1), "indexed": True, "method": "__newindex"})
    return out


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    calls = scan(ref)
    methods = sorted({c["method"] for c in calls})
    dst = os.path.join(ROOT, "tests", "golden", "lua_tensor_methods.json")
    json.dump({"source": "tools/scan_lua_methods.py over the reference's Lua scripts", "methods": methods, "calls": calls}, open(dst, "w"), indent=0)
    print(len(calls), "call sites;", "methods:", " ".join(methods))


if __name__ == "__main__":
    main()
