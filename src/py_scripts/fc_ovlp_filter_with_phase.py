from falcon_unzip_b200.ovlp_filter_with_phase import main
import sys
if __name__ == "__main__":
    main(sys.argv)
