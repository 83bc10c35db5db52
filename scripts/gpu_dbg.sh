#!/bin/bash
# scratch: one quick GPU check (edit per experiment)
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
