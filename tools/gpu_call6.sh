#!/bin/bash
timeout 400 python tools/debug_p8.py 2>&1 | grep "CASE\|Error\|error" 
