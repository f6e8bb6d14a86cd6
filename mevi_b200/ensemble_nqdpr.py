"""CLI with the flags of MEVI/ensemble_nqdpr.py:254-270."""
from argparse import ArgumentParser

from .ensemble import combine_main_nqdpr as combine_main


def main(argv=None):
    parser = ArgumentParser()
    parser.add_argument("--dir_path", type=str, required=True)
    parser.add_argument("--ance_file", type=str, required=True)
    parser.add_argument("--fine_file", type=str, default=None)
    parser.add_argument("--coarse_file", type=str, default=None)
    parser.add_argument("--mapping_file", type=str, default=None)
    parser.add_argument("--alphas", type=str, default="0.4")
    parser.add_argument("--betas", type=str, default="0.03")
    parser.add_argument("--gammas", type=str, default="0.02")
    parser.add_argument("--recall_num", type=str, default="5,20,100")
    parser.add_argument("--ofile", type=str, default=None)
    parser.add_argument("--device", type=str, default="cuda", choices=["cuda", "host"],
                        help="cuda: fusion / ranking / look-up kernels of libmevi_b200.so; host: the reference's python dictionaries")
    parser.add_argument("--noensemble", action="store_true", default=False)
    combine_main(parser.parse_args(argv))


if __name__ == "__main__":
    main()
