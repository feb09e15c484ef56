#!/usr/bin/env bash
# Reference tools/update_git_dir.sh runs `git pull` on every worker; here the tree is pushed from this node instead.
exec python "$(dirname "$0")/cluster.py" sync "$@"
