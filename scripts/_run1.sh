#!/usr/bin/env bash
# scratch: the command of the last ad-hoc GPU visit (see scripts/gpu_round.sh for the round's evidence run)
set -u
bash scripts/gpu_round.sh "${1:-r02}"
