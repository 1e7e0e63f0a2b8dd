import json,re
j=json.load(open('gpurun_out/.last_call.json')); t=j['stdout_tail']
for line in t.splitlines():
    if line.startswith('attn='):
        m=re.search(r'(attn=\w+ splits=\d+).*ms_per_token_step": ([0-9.]+).*avg_launch_ms": ([0-9.]+)', line)
        print(m.groups() if m else line[:300])
    elif 'passed' in line or 'failed' in line or 'Error' in line or 'error' in line: print(line[:300])
print("gpu min left", j.get("gpu_minutes_left"))
