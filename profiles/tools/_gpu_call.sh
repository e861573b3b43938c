TAG=v2n bash profiles/tools/round_profile.sh
