#!/usr/bin/env bash
# Reference tools/remote_script.sh: copy keys and clone the repo on every worker.  Here: write the hosts files and rsync this tree to every node.
set -euo pipefail
here="$(dirname "$0")"
python "$here/cluster.py" get_hosts "$@"
python "$here/cluster.py" sync "$@"
python "$here/cluster.py" run_command "$@" -- "cd ~/ps_pytorch_b200 && bash tools/pre_run.sh"
