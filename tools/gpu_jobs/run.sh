#!/bin/bash
# usage: tools/gpu_jobs/run.sh <timeout_s> <job.sh> [gpus] [VAR=VALUE ...]
# submits a job (the VAR=VALUE words are put in front of the remote command: the local environment does not travel),
# retrying while the pod answers "busy" (exit 3)
T=$1; JOB=$2; G=${3:-1}; shift; shift; shift
ENVS="$*"
for i in 1 2 3 4 5 6 7 8; do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- "$ENVS bash $JOB"; else /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$ENVS bash $JOB"; fi
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3
