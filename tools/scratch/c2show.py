import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
c=d["configs"]["config2_val_set_3000_frames"]
print(sys.argv[1], "value", round(d["value"]), "config2 poses/s", round(c.get("poses_per_s",0)), "ms", c.get("ms_per_pass"), c.get("failed"))
