#!/bin/bash
# (bring-up of pnr_wgrad alone; superseded by tools/r2_native.sh, which it now runs)
exec bash tools/r2_native.sh
