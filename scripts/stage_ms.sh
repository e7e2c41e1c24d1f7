#!/bin/bash
# quick per-stage timing on the GPU box: prints resident / e2e Mpix/s and ms per pair of every stage
python bench.py --steps 3 --warmup 3 --cpu-pairs 0 "$@" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['e2e']['value'],1), d['stages_ms_per_pair'])"
