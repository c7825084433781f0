#!/usr/bin/env python
"""CLI with the reference's flags (model.py:419-500):

python vaxmc_cli.py --pro=35,45 --anti=12,25 --pressure=0.,0.02 --dupl_time=5,7 \
    --init_stock=1,1.2 --max_delivery=5,10 --date_range=2020-12-15,2022-07-1
"""
import sys

import vaxmc_b200

if __name__ == "__main__":
    sys.exit(vaxmc_b200.main())
