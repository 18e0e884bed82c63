#!/usr/bin/env bash
# Single-box smoke test of the mesh example: 4 processes on localhost ports 9100-9103.
set -e
cd "$(dirname "$0")/../../.."
pids=()
for i in 0 1 2 3; do
  python examples/p2p/remote_tcp/mesh_client.py --node-id $i &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
echo "mesh smoke test finished"
