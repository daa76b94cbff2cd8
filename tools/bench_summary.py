import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"],1), round(d["ms_per_step"],3), round(d["e2e"]["value"],1))
        print({k:round(v,3) for k,v in d["stage_ms_per_step"].items()})
    except Exception as e:
        print(f, "ERR", e)
