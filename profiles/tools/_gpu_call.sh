TAG=v2m bash profiles/tools/round_profile.sh
