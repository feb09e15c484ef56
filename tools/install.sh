#!/usr/bin/env bash
# Reference tools/install.sh installs pdsh for fan-out ssh.  cluster.py does its own fan-out; this only checks ssh + rsync.
for t in ssh rsync; do command -v $t >/dev/null || { echo "missing: $t"; exit 1; }; done; echo "ssh and rsync present"
