#!/usr/bin/env bash
# Reference tools/killall.sh runs `killall python` on every worker.  Here only the jobs started by `cluster.py launch` are stopped,
# by the process groups recorded in their pid files — never by name.
exec python "$(dirname "$0")/cluster.py" kill_all "$@"
