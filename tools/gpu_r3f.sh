#!/bin/bash
timeout 300 python -m pytest tests/test_table_match.py -m gpu -q --timeout 200 -p no:cacheprovider 2>&1 | tail -12
