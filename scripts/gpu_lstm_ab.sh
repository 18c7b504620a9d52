#!/bin/bash
PTRAIN=0 python scripts/prof_lstm_steps.py 2>&1 | grep -A10 "^forward B"
