import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value",d["value"],"ms",d["ms_per_step"],"per_rank",d.get("per_rank_ms"))
print("cfg3",d["cfg3"]["value"],d["cfg3"]["ms_per_step"],d["cfg3"]["level_hist"])
print("e2e",d["e2e"]["value"],d["e2e"]["ms_per_step"],d["e2e"].get("link_bound_ms"))
