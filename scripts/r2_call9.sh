#!/usr/bin/env bash
cd "$(dirname "$0")/.."
timeout 300 python scripts/dbg_explicit.py 2>&1 | grep -v Warn | tail -20
