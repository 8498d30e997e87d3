import json,sys
d=json.loads(sys.stdin.read()); print(d["config"]["workload"], "%.3e"%d["value"], "ms/step %.3f"%d["ms_per_step"], "kernel %.3f"%d["roofline"]["kernel_ms"])
