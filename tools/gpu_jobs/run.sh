#!/bin/bash
# usage: tools/gpu_jobs/run.sh <timeout_s> <job.sh> [gpus]  — submits a job, retrying while the pod answers "busy" (exit 3)
T=$1; JOB=$2; G=${3:-1}
for i in 1 2 3 4 5 6 7 8; do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- "bash $JOB"; else /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "bash $JOB"; fi
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3
