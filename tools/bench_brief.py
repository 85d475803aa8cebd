"""Print the key figures of bench.py JSON lines: python tools/bench_brief.py file.json ..."""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as ex:  # noqa: BLE001
        print(f, "unreadable:", ex)
        continue
    r = d.get("roofline") or {}
    e = d.get("e2e") or {}
    print(f"{f}: value {d.get('value')} ms/step {d.get('ms_per_step')} e2e {e.get('value')} u8host {e.get('u8_host_api_value')} "
          f"frac {r.get('frac')} hot_ms {r.get('launch_ms')} dirs/launch {r.get('directions_per_launch')} launches {d.get('gpu_launches')}")
