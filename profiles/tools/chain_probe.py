import os, sys, json
sys.path.insert(0, "/root/repo")
import circuitsimulator_b200 as bg
S = "/root/repo/tests/golden/streams"
meta = json.load(open(os.path.join(S, "meta.json")))["phase_estimation_chain"]["streams"]
exact = {"0": 0.9267766952966369, "00": 0.21338834764831827, "010": 0.5, "000": 0.0, "10": 0.03661165235168153, "100": 0.0, "110": 0.0}
for key, ent in sorted(meta.items()):
    txt = open(os.path.join(S, ent["stream"])).read().split()
    for k in (8, 12):
        txt[6] = str(k)
        vals = []
        for seed in range(3):
            num, den, _ = bg.run_backend("\n".join(txt) + "\n", env={"BG_SEED": seed})
            vals.append(0.0 if num == 0 else 2.0 ** ent["v_minus_u"] * num / den)
        print(key, "k", k, ["%.4f" % v for v in vals], "exact", exact.get(key))
