#!/bin/bash
for cfg in "10000 10000 200" "5000 7777 150" "3000 4100 256" "700 900 64"; do echo "== $cfg"; timeout 600 python scripts/ts_probe.py $cfg 2>&1 | tail -n 5; done
