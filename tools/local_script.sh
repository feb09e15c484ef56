#!/usr/bin/env bash
# Reference tools/local_script.sh <master address>: push the repo, hosts files and ssh config to the master node.
set -euo pipefail
addr="${1:?usage: local_script.sh <head node address> [ssh user]}"; user="${2:-$USER}"; here="$(cd "$(dirname "$0")" && pwd)"
rsync -az --exclude .git --exclude gpurun_out -e "ssh -F $here/config" "$here/.." "$user@$addr:~/ps_pytorch_b200"
echo "repository copied to $user@$addr:~/ps_pytorch_b200 — next: ssh there and run tools/remote_script.sh"
