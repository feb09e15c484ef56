#!/bin/bash
# One process prepares every dataset before the job starts (reference src/data_prepare.sh).
cd "$(dirname "$0")"
python data/data_prepare.py "$@"
