#!/bin/bash
# usage: tools/run_guarded.sh SECONDS command...   - runs the command in its own session; at the deadline the whole
# process group gets SIGTERM, five seconds later SIGKILL (torchrun and every rank), so that a hung multi-GPU run cannot
# keep the box busy until the outer limit.  Exit code 124 on timeout.
limit=$1; shift
setsid "$@" &
pid=$!
( sleep "$limit"; kill -TERM -- -"$pid" 2>/dev/null; sleep 5; kill -KILL -- -"$pid" 2>/dev/null ) &
watch=$!
wait "$pid"; rc=$?
if kill -0 "$watch" 2>/dev/null; then kill "$watch" 2>/dev/null; else rc=124; fi
exit $rc
