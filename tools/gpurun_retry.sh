#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3: nothing charged).  Usage: tools/gpurun_retry.sh [gpurun options] -- 'command'
for attempt in $(seq 1 20); do
	/usr/local/graft/bin/gpurun "$@"
	rc=$?
	if [ $rc -ne 3 ]; then exit $rc; fi
	sleep 45
done
exit 3
