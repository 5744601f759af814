#!/bin/bash
P=tools/probes/ring_probe
for f in 0 1 2 4 8 10 11 15; do $P 6 16 4 2560 64 0 1 3 0 $f; done
for f in 0 2 10 11 15; do $P 6 16 4 2560 64 0 1 3 2 $f; done
for f in 0 10; do $P 6 16 4 2560 0 0 1 3 2 $f; done
