int dirac_oracle_placeholder(void){return 0;}
