TAG=v2q bash profiles/tools/round_profile.sh
